timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn" -p no:cacheprovider 2>&1 | tail -4
PROFILE_TIME=1 timeout 300 python tools/profile_kernels.py attn 2>&1 | tail -3
timeout 300 python tools/attn_timeline.py 2>&1 | grep -E "iter [345] "
