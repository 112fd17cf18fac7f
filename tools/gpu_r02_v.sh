mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn or small_ops or cast or norm" -p no:cacheprovider 2>&1 | tail -3
for pr in 0 1; do
VIDCHAP_SIDE_PRIORITY=$pr timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio $pr step', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['loss'])"
done
python tools/graph_timeline.py 2>&1 | tail -48
rm -f gpurun_out/graph_timeline.trace.json
