timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_e2e_more_gpu.py -m gpu -q -x -k "generate or greedy or beam or decode" -p no:cacheprovider 2>&1 | tail -3
python tools/time_generate.py 2>&1 | tail -4
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 > gpurun_out/r02_bench_z_decode.json 2> gpurun_out/r02_bench_z_decode.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_decode.json').read().strip().splitlines()[-1]); print('decode', d['value'], d['ms_per_step'], d['decode_loop'], d['roofline']['frac'], d['e2e']['value'])"
