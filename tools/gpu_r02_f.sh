mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_f_base.json 2> gpurun_out/r02_bench_f_base.err; tail -c 900 gpurun_out/r02_bench_f_base.json; tail -3 gpurun_out/r02_bench_f_base.err
timeout 600 python bench.py --model t5-large --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_f_large.json 2> gpurun_out/r02_bench_f_large.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_f_large.json').read().strip().splitlines()[-1]); print('t5-large', d['ms_per_step'], d['value'], d['step_tensor_frac'], d['roofline']['frac'])" || tail -5 gpurun_out/r02_bench_f_large.err
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 > gpurun_out/r02_bench_f_decode.json 2> gpurun_out/r02_bench_f_decode.err; tail -c 1500 gpurun_out/r02_bench_f_decode.json; tail -3 gpurun_out/r02_bench_f_decode.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-900
