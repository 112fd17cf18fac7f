python tools/time_gemm_epi.py
VIDCHAP_GEMM_DBG=1 python tools/time_gemm_epi.py
VIDCHAP_GEMM_DBG=2 python tools/time_gemm_epi.py
