mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02_gputest_e.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r02_gputest_e.log
grep -E "FAILED|Error" gpurun_out/r02_gputest_e.log | cut -c1-300 | head -20
PROFILE_TIME=1 timeout 300 python tools/profile_kernels.py attn 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 1500 gpurun_out/r02_bench_e.json
