for k in 1 2; do
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_z_decode.json 2> gpurun_out/r02_bench_z_decode.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_decode.json').read().strip().splitlines()[-1]); print('decode', d['value'], d['ms_per_step'], d['decode_loop'], d['roofline']['frac'], d['e2e']['value'])"
done
