# Round-2 GPU call A: microbench + all GPU tests (verbose parity numbers) + A/B of the engine switches + attention baseline counters.
mkdir -p gpurun_out
nvidia-smi -L
./tools/microbench/tmem_bw > gpurun_out/r02_tmem_bw.txt 2>&1; cat gpurun_out/r02_tmem_bw.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02_gputest_a.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r02_gputest_a.log
grep -E "^\[|FAILED|passed|failed|Error" gpurun_out/r02_gputest_a.log | cut -c1-300 | tail -60
i=0
for variant in "X=1" "VIDCHAP_FUSE_CROSS_KV=1" "VIDCHAP_WGRAD_STREAM=1" "VIDCHAP_FUSE_CROSS_KV=1 VIDCHAP_WGRAD_STREAM=1"; do
  env $variant timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a_$i.json 2> gpurun_out/r02_bench_a_$i.err
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/r02_bench_a_$i.json').read().strip().splitlines()[-1])
    print('$variant', 'ms', round(d['ms_per_step'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'gemm frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
except Exception as e:
    print('$variant', 'FAILED', e); print(open('gpurun_out/r02_bench_a_$i.err').read()[-1500:])
"
  i=$((i+1))
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:attn --csv --log-file gpurun_out/r02_attn_counters_a.csv python tools/profile_kernels.py attn > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_attn_counters_a.csv | tail -24 | cut -c1-260
