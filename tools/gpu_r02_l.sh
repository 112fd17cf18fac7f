mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "decode" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -p no:cacheprovider -k "generate" 2>&1 | tail -6
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02_bench_l_decode.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('decode', d['ms_per_step'], d['value'], d['decode_loop'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])" || tail -5 gpurun_out/r02_bench_l_decode.err
