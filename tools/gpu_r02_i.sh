mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_profile.csv python tools/one_step.py > gpurun_out/r02_step_profile.log 2>&1
tail -2 gpurun_out/r02_step_profile.log
python tools/summarize_step_profile.py gpurun_out/r02_step_profile.csv gpurun_out/r02_step
cat gpurun_out/r02_step_tensor_pipe.json gpurun_out/r02_step_gemm_traffic.json
