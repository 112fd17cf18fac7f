timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_e2e_gpu.py -m gpu -q -x -k "decode or generate or greedy or beam" -p no:cacheprovider 2>&1 | tail -3
python tools/decode_timeline.py 2>&1 | grep -v Warn | tail -13
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 > gpurun_out/r02_bench_z_decode.json 2> gpurun_out/r02_bench_z_decode.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_decode.json').read().strip().splitlines()[-1]); print('decode', d['value'], d['ms_per_step'], d['decode_loop'], d['roofline']['frac'], d['e2e']['value'])"
