mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn" -p no:cacheprovider 2>&1 | tail -15
PROFILE_TIME=1 timeout 300 python tools/profile_kernels.py attn 2>&1 | tail -4
PROFILE_TIME=1 PROFILE_QLK=0 timeout 300 python tools/profile_kernels.py attn 2>&1 | tail -3
PROFILE_TIME=1 VIDCHAP_ATTN_FWD_PAIR=0 timeout 300 python tools/profile_kernels.py attn 2>&1 | tail -3
