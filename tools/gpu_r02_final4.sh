timeout 900 python -m pytest tests/test_e2e_more_gpu.py -m gpu -q -x -p no:cacheprovider -k "second_device" 2>&1 | tail -8
