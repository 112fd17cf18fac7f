"""Launches the hot kernels at config-2 shapes a few times (driver for `ncu --set full -k regex:<name>`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vidchapters_b200.ops import CudaOps
from vidchapters_b200.engine import relative_position_bucket

ops = CudaOps()
dev = "cuda"
g = torch.Generator().manual_seed(0)
B, H, L = 16, 12, 1000
inner = H * 64
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "attn"):
    qkv = (torch.randn(B * L, 3 * inner, generator=g) * 0.5).to(dev).bfloat16()
    lut = relative_position_bucket(torch.arange(2 * L - 1) - (L - 1), True).to(torch.int32).to(dev)
    bias = torch.randn(32, H, generator=g).to(dev)[lut.long()].t().contiguous()   # T5 bucketed bias, as in the train step
    lens = torch.randint(L // 2, L + 1, (B,), generator=g)                        # bench.py's ragged ASR lengths
    kmask = (torch.arange(L)[None] < lens[:, None]).to(torch.uint8).to(dev)
    out = torch.zeros(B * L, inner, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    kw = dict(q_col=0, k_col=inner, v_col=2 * inner, B=B, H=H, Lq=L, Lk=L, bias_rel=bias, kmask=kmask, causal=False, scale=1.0)
    dout = (torch.randn(B * L, inner, generator=g) * 0.5).to(dev).bfloat16()
    delta = torch.zeros(B, H, L, device=dev); dq = torch.zeros(B * L, inner, device=dev)
    dqkv = torch.zeros(B * L, 3 * inner, device=dev, dtype=torch.bfloat16)
    db = torch.zeros(H, 2 * L - 1, device=dev)
    drop = (0xC0FFEE, 6554) if os.environ.get("PROFILE_DROPOUT", "1") == "1" else (0, 0)   # reference default p = 0.1
    kw["drop"] = drop
    kw["q_like_k"] = os.environ.get("PROFILE_QLK", "1") == "1"     # as the train step calls the text encoder's attention
    for _ in range(3):
        ops.attn_fwd(qkv, qkv, qkv, out=out, lse2=lse, **kw)
        ops.attn_bwd(qkv, qkv, qkv, out=out, lse2=lse, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dqkv, dk_col=inner,
                     dv=dqkv, dv_col=2 * inner, dbias_rel=db, bucket_lut=lut, **kw)
    if os.environ.get("PROFILE_TIME", "0") == "1":
        for name, fn in (("attn_fwd", lambda: ops.attn_fwd(qkv, qkv, qkv, out=out, lse2=lse, **kw)),
                         ("attn_bwd", lambda: ops.attn_bwd(qkv, qkv, qkv, out=out, lse2=lse, dout=dout, do_col=0, delta=delta,
                                                           dq_acc=dq, dk=dqkv, dk_col=inner, dv=dqkv, dv_col=2 * inner,
                                                           dbias_rel=db, bucket_lut=lut, **kw))):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (encoder shape B=16 H=12 L=1000, q_like_k={kw['q_like_k']})")
if which in ("all", "gemm"):
    M, N, K = 16000, 3072, 768
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.5).to(dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    dW = torch.zeros(N, K, device=dev)
    dA = torch.zeros(M, K, device=dev)
    for _ in range(3):
        ops.gemm(A, W, C, act=1)                                    # forward wi + relu
        ops.gemm(C, W, dA, b_mn=True)                               # dgrad
        ops.gemm(C, A, dW, a_mn=True, b_mn=True, atomic=True, splits=5)  # wgrad
if which in ("gemm_wi", "gemm_resid"):
    from vidchapters_b200.ops import drop_spec
    M = 16000
    N, K = (3072, 768) if which == "gemm_wi" else (768, 768)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    if which == "gemm_wi":
        C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); kw = dict(act=1, drop=drop_spec(0.1, 7))
    else:
        C = torch.zeros(M, N, device=dev); kw = dict(residual=torch.randn(M, N, device=dev), drop=drop_spec(0.1, 7))
    for _ in range(4):
        ops.gemm(A, W, C, **kw)
torch.cuda.synchronize()
print("done")
