timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_e2e_gpu.py -m gpu -q -x -p no:cacheprovider -k "small_ops or golden or switches or floor or tier" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['loss'])"
python tools/graph_timeline.py 2>&1 | grep -E "bias_fold|span"
rm -f gpurun_out/graph_timeline.trace.json
