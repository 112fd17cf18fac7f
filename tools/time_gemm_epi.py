"""Decompose the GEMM time at the train-step shapes: K sweep (fixed cost per tile vs mainloop) and, with
VIDCHAP_GEMM_DBG=1 / 2 (set per process), the kernel without its epilogue body / without its stores.
python tools/time_gemm_epi.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidchapters_b200.ops import CudaOps, drop_spec
ops = CudaOps()
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)

def t(fn, n=8):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:]); return ts[len(ts) // 2]

M = 16000
print("dbg =", os.environ.get("VIDCHAP_GEMM_DBG", "0"))
for N, Ks, kw_name in ((3072, (768, 1536, 3072), "relu+drop bf16"), (3072, (768,), "plain bf16"), (2304, (768,), "plain bf16"),
                       (768, (768, 1536, 3072), "resid f32"), (768, (768, 3072), "plain f32"), (768, (768, 3072), "plain bf16")):
    for K in Ks:
        A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
        W = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
        if kw_name == "relu+drop bf16":
            C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); kw = dict(act=1, drop=drop_spec(0.1, 7))
        elif kw_name == "plain bf16":
            C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); kw = {}
        elif kw_name == "plain f32":
            C = torch.zeros(M, N, device=dev); kw = {}
        else:
            C = torch.zeros(M, N, device=dev); R = torch.randn(M, N, device=dev); kw = dict(residual=R, drop=drop_spec(0.1, 7))
        us = t(lambda: ops.gemm(A, W, C, **kw))
        print(f"M={M} N={N} K={K} {kw_name:>15}: {us:7.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s")
M = 4096
for N, K, kw_name in ((768, 768, "resid f32"), (768, 3072, "resid f32"), (2304, 768, "plain bf16"), (3072, 768, "relu+drop bf16")):
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    if kw_name == "resid f32":
        C = torch.zeros(M, N, device=dev); R = torch.randn(M, N, device=dev); kw = dict(residual=R, drop=drop_spec(0.1, 7))
    elif kw_name == "plain bf16":
        C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); kw = {}
    else:
        C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); kw = dict(act=1, drop=drop_spec(0.1, 7))
    us = t(lambda: ops.gemm(A, W, C, **kw))
    print(f"M={M} N={N} K={K} {kw_name:>15}: {us:7.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s")
