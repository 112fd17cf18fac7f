mkdir -p gpurun_out
python tools/graph_timeline.py 2>&1 | tail -60
ls -la gpurun_out/graph_timeline*; rm -f gpurun_out/graph_timeline.trace.json
