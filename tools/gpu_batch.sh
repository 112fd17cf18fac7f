# One gpurun call = tests, benches and profiles (each call costs ~5 min of box overhead, so batch).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/t_ops.log 2>&1 || { tail -40 gpurun_out/t_ops.log; exit 1; }
tail -2 gpurun_out/t_ops.log
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; grep -E "passed|failed|^E |config2 full" gpurun_out/t_e2e.log | head -20
J='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_eager"]["ms_per_step"], d["roofline"]["achieved"], d["loss"])'
echo "=== BENCH new split-K heuristic"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-gemms gpurun_out/gemm_shapes_r01g.txt 2>&1 | tail -1 | python -c "$J"
echo "=== BENCH old split-K heuristic"
VIDCHAP_SPLITS_OLD=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$J"
echo done
