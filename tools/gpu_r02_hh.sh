timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -k "switches" -p no:cacheprovider 2>&1 | tail -3
for f in 0 1 0 1; do
VIDCHAP_DEC_WGRAD_STREAM=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dec wgrad stream $f: step', d['ms_per_step'], d['value'], 'loss', d['loss'], d['clocks']['sm_mhz'])"
done
