# One gpurun call = tests, benches and profiles (each call costs ~5 min of box overhead, so batch).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/t_ops.log 2>&1 || { tail -40 gpurun_out/t_ops.log; exit 1; }
tail -2 gpurun_out/t_ops.log
timeout 600 python -m pytest tests/test_e2e_gpu.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; grep -E "passed|failed|^E " gpurun_out/t_e2e.log | head -20
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
J='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_eager"]["ms_per_step"], d["roofline"]["achieved"], d["loss"], d.get("cpu_baseline"))'
echo "=== BENCH default flags"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | python -c "$J"
echo "=== BENCH 20 steps, gemm dump"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-gemms gpurun_out/gemm_shapes_r01f.txt 2>&1 | tail -1 | python -c "$J"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r01j.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 2 -o gpurun_out/prof_gemm2_r01f python tools/profile_kernels.py gemm > gpurun_out/p9.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 1 -c 1 -o gpurun_out/prof_attn_bwd_r01f python tools/profile_kernels.py attn > gpurun_out/p7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/prof_attn_fwd_r01f python tools/profile_kernels.py attn > gpurun_out/p8.log 2>&1
echo done
