for f in 0 1 0 1 0 1; do
VIDCHAP_DEC_WGRAD_STREAM=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dec wgrad stream $f: step', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
