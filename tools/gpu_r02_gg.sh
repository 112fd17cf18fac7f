for k in 1 2 3; do
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_z_decode.json 2> gpurun_out/r02_bench_z_decode.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_decode.json').read().strip().splitlines()[-1]); print('decode', d['value'], d['ms_per_step'], d['decode_loop']['tokens_per_sec'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
