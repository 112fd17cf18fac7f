# One gpurun call = tests, A/B benches and profiles (each call costs ~5 min of box overhead, so batch).
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/t_ops.log 2>&1 || { tail -40 gpurun_out/t_ops.log; exit 1; }
tail -2 gpurun_out/t_ops.log
timeout 600 python -m pytest tests/test_e2e_gpu.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; grep -E "^\[|passed|failed|^E " gpurun_out/t_e2e.log | head -40
J='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["loss"])'
echo "=== BENCH default (16 bwd warps)"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms gpurun_out/gemm_shapes_r01e.txt 2>&1 | tail -1 | python -c "$J"
echo "=== BENCH 8 bwd warps"
VIDCHAP_ATTN_BWD_WARPS=8 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$J"
echo "=== BENCH dropout 0"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dropout 0 2>&1 | tail -1 | python -c "$J"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 1 -c 1 -o gpurun_out/prof_attn_bwd_r01e python tools/profile_kernels.py attn > gpurun_out/p7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/prof_attn_fwd_r01e python tools/profile_kernels.py attn > gpurun_out/p8.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
echo done
