mkdir -p gpurun_out
for k in attn_fwd_pair attn_bwd_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02n_$k python tools/profile_kernels.py attn > gpurun_out/r02n_$k.log 2>&1
  ls -la gpurun_out/r02n_$k.ncu-rep
done
