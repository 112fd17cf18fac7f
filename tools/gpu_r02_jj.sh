python tools/graph_timeline.py --model t5-large --out gpurun_out/graph_timeline_large 2>&1 | grep -v Warn | tail -34
rm -f gpurun_out/graph_timeline_large.trace.json
