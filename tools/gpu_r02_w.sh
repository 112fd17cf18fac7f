mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['loss'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'])"
python tools/graph_timeline.py 2>&1 | tail -42
rm -f gpurun_out/graph_timeline.trace.json
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
