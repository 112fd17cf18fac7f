mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_profile.csv python tools/one_step.py > gpurun_out/r02_step_profile.log 2>&1
python tools/summarize_step_profile.py gpurun_out/r02_step_profile.csv gpurun_out/r02_step | head -30
python tools/graph_timeline.py 2>&1 | tail -45
rm -f gpurun_out/graph_timeline.trace.json
python tools/attn_timeline.py > gpurun_out/r02_attn_bwd_timeline.txt 2>&1; tail -5 gpurun_out/r02_attn_bwd_timeline.txt
python tools/time_gemm_epi.py > gpurun_out/r02_gemm_epi_times.txt 2>&1; tail -3 gpurun_out/r02_gemm_epi_times.txt
python tools/time_norm_bwd.py > gpurun_out/r02_norm_bwd_times.txt 2>&1; cat gpurun_out/r02_norm_bwd_times.txt
python tools/time_attn_shapes.py > gpurun_out/r02_attn_shape_times.txt 2>&1; cat gpurun_out/r02_attn_shape_times.txt
