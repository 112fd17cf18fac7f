mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "fused_lm_head or cross_entropy or gemm" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_e2e_more_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused CE   ', d['ms_per_step'], d['value'], d['loss'])"
VIDCHAP_FUSED_CE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unfused CE ', d['ms_per_step'], d['value'], d['loss'])"
