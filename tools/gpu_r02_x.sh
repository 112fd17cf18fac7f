timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn" -p no:cacheprovider 2>&1 | tail -3
python tools/time_attn_shapes.py
echo "--- baseline library"
cp vidchapters_b200/libvidchap_base.so vidchapters_b200/libvidchap.so
python tools/time_attn_shapes.py
