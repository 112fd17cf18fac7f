# One gpurun call = tests, A/B benches and profiles (each call costs ~5 min of box overhead, so batch).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/t_ops.log 2>&1 || { tail -40 gpurun_out/t_ops.log; exit 1; }
tail -2 gpurun_out/t_ops.log
timeout 600 python -m pytest tests/test_e2e_gpu.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; grep -E "passed|failed|^E " gpurun_out/t_e2e.log | head -20
J='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_eager"]["ms_per_step"], d["roofline"]["achieved"], d["loss"])'
echo "=== BENCH default (PDL on)"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$J"
echo "=== BENCH PDL off"
VIDCHAP_PDL=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$J"
echo "=== BENCH default again, 20 steps"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$J"
echo done
