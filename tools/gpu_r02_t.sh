timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm or fused_lm" -p no:cacheprovider 2>&1 | tail -3
python tools/time_gemm_epi.py
