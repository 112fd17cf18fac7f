# ncu full captures (with source counters) of the two attention kernels at config-2 encoder shapes
mkdir -p gpurun_out
for k in attn_fwd attn_bwd_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02b_$k python tools/profile_kernels.py attn > gpurun_out/r02b_$k.log 2>&1; tail -3 gpurun_out/r02b_$k.log
  ls -la gpurun_out/r02b_$k.ncu-rep
done
