mkdir -p gpurun_out
for k in gemm_wi gemm_resid; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 2 -c 1 -f -o gpurun_out/r02s_$k python tools/profile_kernels.py $k > gpurun_out/r02s_$k.log 2>&1
  ls -la gpurun_out/r02s_$k.ncu-rep
done
