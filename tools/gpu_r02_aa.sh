timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn" -p no:cacheprovider 2>&1 | tail -3
python tools/time_attn_shapes.py
python tools/attn_timeline.py | grep -E "iter 4|iter 5" | head -24
