mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm or fused_lm" -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms gpurun_out/r02_gemm_shapes_p.txt 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], d['value'], 'gemm frac', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['loss'])"
head -14 gpurun_out/r02_gemm_shapes_p.txt
