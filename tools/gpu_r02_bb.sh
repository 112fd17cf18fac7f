timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "presets" -p no:cacheprovider 2>&1 | tail -3
python tools/decode_timeline.py 2>&1 | tail -20
cat gpurun_out/decode_timeline.kernels.txt | head -24
