timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "norm" -p no:cacheprovider 2>&1 | tail -3
echo "--- old kernel"; VIDCHAP_NORM_BWD_WIDE=0 python tools/time_norm_bwd.py
echo "--- wide occ3"; python tools/time_norm_bwd.py
echo "--- wide occ2"; VIDCHAP_NORM_BWD_OCC=2 python tools/time_norm_bwd.py
