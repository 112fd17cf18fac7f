timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_e2e_gpu.py -m gpu -q -x -k "decode or generate or greedy or beam" -p no:cacheprovider 2>&1 | tail -3
python tools/decode_timeline.py 2>&1 | grep -v Warn | tail -12
VIDCHAP_DECODE_SPLITK=0 python tools/decode_timeline.py 2>&1 | grep -E "span|48, 1" 
