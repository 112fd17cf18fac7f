"""Per-kernel parity on the GPU: every C-ABI entry point vs the plain-torch restatement (oracle/torch_ops.py)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def gen(seed=0):
    return torch.Generator(device="cpu").manual_seed(seed)


# bf16 output rounding alone is ~1.7e-3 rel-L2; fp32 outputs should be ~1e-6.
BF16_TOL, F32_TOL = 4e-3, 2e-5


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,tile_n", [
    (128, 64, 64, 0, 0, 64), (256, 256, 256, 0, 0, 256), (1600, 776, 768, 0, 0, 0), (1000, 2304, 768, 0, 0, 0),
    (300, 1100, 768, 0, 0, 0), (130, 2004, 128, 0, 0, 0),  # ragged N: not a multiple of 8 (vocab 32100-style)
    (256, 256, 256, 0, 1, 256), (1001, 768, 3072, 0, 1, 0), (128, 128, 128, 1, 0, 128), (384, 512, 1000, 1, 1, 0),
])
def test_gemm_layouts(cuda_ops, torch_ops, M, N, K, a_mn, b_mn, tile_n):
    g = gen(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).to(DEV).bfloat16()
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    for dt, tol in ((torch.float32, F32_TOL), (torch.bfloat16, BF16_TOL)):
        Np = (N + 7) // 8 * 8
        out = torch.zeros(M, Np, device=DEV, dtype=dt)[:, :N]
        ref = torch.zeros(M, Np, device=DEV, dtype=dt)[:, :N]
        cuda_ops.gemm(A_st, B_st, out, a_mn=bool(a_mn), b_mn=bool(b_mn), tile_n=tile_n)
        torch_ops.gemm(A_st, B_st, ref, a_mn=bool(a_mn), b_mn=bool(b_mn))
        assert rel(out, ref) < tol


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,tile_n", [
    (2048, 1024, 512, 0, 0, 256), (2048, 1024, 512, 0, 1, 256), (2048, 1024, 512, 1, 0, 256), (2048, 1024, 512, 1, 1, 256),
    (1000, 776, 320, 0, 0, 128), (1000, 776, 320, 0, 1, 128), (1000, 776, 320, 1, 1, 128), (16000, 3072, 768, 0, 0, 0),
    (4096, 32200, 768, 0, 0, 0), (3072, 768, 16000, 1, 1, 0),
])
def test_gemm_cta_pair_kernel(cuda_ops, torch_ops, M, N, K, a_mn, b_mn, tile_n):
    """Shapes / forced tiles that take the 2-CTA (cta_group::2) kernel when VIDCHAP_GEMM_PAIR != 0."""
    g = gen(M + N + K + 7)
    A = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).to(DEV).bfloat16()
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    bias = torch.randn(N, generator=g).to(DEV)
    out = torch.zeros(M, N, device=DEV)
    ref = torch.zeros(M, N, device=DEV)
    splits = 3 if (a_mn and b_mn) else 1   # wgrad-style: split-K with fp32 atomics (no bias: every split would add it)
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), bias=None if splits > 1 else bias, atomic=splits > 1, splits=splits)
    cuda_ops.gemm(A_st, B_st, out, tile_n=tile_n, **kw)
    torch_ops.gemm(A_st, B_st, ref, **kw)
    assert rel(out, ref) < F32_TOL


@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_gemm_epilogues(cuda_ops, torch_ops, act):
    M, N, K = 520, 2048, 768
    g = gen(act)
    A = (torch.randn(M, K, generator=g) * 0.3).to(DEV).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.3).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    resid = torch.randn(M, N, generator=g).to(DEV)
    aux = torch.randn(M, N, generator=g).to(DEV).bfloat16()
    adev = torch.tensor([0.7], device=DEV)
    kw = dict(bias=bias, residual=resid, act=act, aux=aux if act >= 3 else None, alpha=0.5, alpha_dev=adev)
    out = torch.empty(M, N, device=DEV)
    ref = torch.empty(M, N, device=DEV)
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16) if act == 2 else None
    pre_ref = torch.empty_like(pre) if act == 2 else None
    cuda_ops.gemm(A, B, out, pre_out=pre, **kw)
    torch_ops.gemm(A, B, ref, pre_out=pre_ref, **kw)
    assert rel(out, ref) < 1e-4
    if act == 2:
        assert rel(pre, pre_ref) < BF16_TOL


def test_gemm_splitk_atomic(cuda_ops, torch_ops):
    M, N, K = 768, 3072, 4096  # wgrad-shaped: dW[N_out=768... reduction over tokens
    g = gen(5)
    A = (torch.randn(K, M, generator=g) * 0.3).to(DEV).bfloat16()
    B = (torch.randn(K, N, generator=g) * 0.3).to(DEV).bfloat16()
    out = torch.ones(M, N, device=DEV)
    ref = torch.ones(M, N, device=DEV)
    cuda_ops.gemm(A, B, out, a_mn=True, b_mn=True, atomic=True, splits=7)
    torch_ops.gemm(A, B, ref, a_mn=True, b_mn=True, atomic=True)
    assert rel(out, ref) < F32_TOL


def _attn_case(B, H, Lq, Lk, causal, with_bias, with_mask, scale, self_attn, seed=0):
    g = gen(seed)
    inner = H * 64
    if self_attn:
        qkv = (torch.randn(B * Lq, 3 * inner, generator=g) * 0.5).to(DEV).bfloat16()
        q = k = v = qkv
        cols = dict(q_col=0, k_col=inner, v_col=2 * inner)
    else:
        q = (torch.randn(B * Lq, inner, generator=g) * 0.5).to(DEV).bfloat16()
        k = v = (torch.randn(B * Lk, 2 * inner, generator=g) * 0.5).to(DEV).bfloat16()
        cols = dict(q_col=0, k_col=0, v_col=inner)
    bias = (torch.randn(H, Lq + Lk - 1, generator=g)).to(DEV) if with_bias else None
    if with_bias == "t5":   # bucketed like T5: constant beyond +-128 positions -> uniform tiles take the kernels' fast path
        bias = torch.randn(33, H, generator=g).to(DEV)[_t5ish_lut(Lq, Lk).long()].t().contiguous()
    kmask = None
    if with_mask:
        lens = torch.randint(max(1, Lk // 2), Lk + 1, (B,), generator=g)
        kmask = (torch.arange(Lk)[None, :] < lens[:, None]).to(torch.uint8).to(DEV)
    return q, k, v, cols, bias, kmask


def _t5ish_lut(Lq, Lk):
    return (((torch.arange(Lq + Lk - 1) - (Lq - 1)).clamp(-128, 128) + 128) // 8).to(torch.int32).to(DEV)


ATTN_CASES = [
    # B,H,Lq,Lk,causal,bias,mask,scale,self
    (2, 12, 100, 100, False, False, False, 0.125, True),     # ViT
    (2, 12, 1000, 1000, False, True, True, 1.0, True),       # T5 encoder
    (2, 12, 256, 256, True, True, True, 1.0, True),          # T5 decoder self
    (2, 12, 256, 1100, False, False, True, 1.0, False),      # cross
    (1, 4, 37, 130, False, True, True, 1.0, False),          # ragged
    (1, 2, 10, 10, False, False, False, 0.125, True),        # tiny (config 1 ViT)
    (1, 2, 32, 32, True, True, False, 1.0, True),
    (2, 4, 640, 640, False, "t5", False, 1.0, True),         # bucketed bias, all keys attend: constant-bias fast tiles
    (2, 4, 600, 600, False, "t5", True, 1.0, True),          # + ragged tail and key masks
    (1, 4, 500, 500, True, "t5", False, 1.0, True),          # + causal
]


@pytest.mark.parametrize("case", ATTN_CASES)
def test_attn_fwd(cuda_ops, torch_ops, case):
    B, H, Lq, Lk, causal, wb, wm, scale, sa = case
    q, k, v, cols, bias, kmask = _attn_case(B, H, Lq, Lk, causal, wb, wm, scale, sa)
    out = torch.zeros(B * Lq, H * 64, device=DEV, dtype=torch.bfloat16)
    ref = torch.zeros_like(out)
    lse = torch.zeros(B, H, Lq, device=DEV)
    lse_ref = torch.zeros_like(lse)
    kw = dict(B=B, H=H, Lq=Lq, Lk=Lk, bias_rel=bias, kmask=kmask, causal=causal, scale=scale, **cols)
    cuda_ops.attn_fwd(q, k, v, out=out, lse2=lse, **kw)
    torch_ops.attn_fwd(q, k, v, out=ref, lse2=lse_ref, **kw)
    torch.cuda.synchronize()
    assert rel(out, ref) < 8e-3, rel(out, ref)
    assert (lse - lse_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("case", ATTN_CASES)
def test_attn_bwd(cuda_ops, torch_ops, case):
    B, H, Lq, Lk, causal, wb, wm, scale, sa = case
    q, k, v, cols, bias, kmask = _attn_case(B, H, Lq, Lk, causal, wb, wm, scale, sa, seed=3)
    inner = H * 64
    g = gen(11)
    dout = (torch.randn(B * Lq, inner, generator=g) * 0.5).to(DEV).bfloat16()
    out = torch.zeros(B * Lq, inner, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Lq, device=DEV)
    kw = dict(B=B, H=H, Lq=Lq, Lk=Lk, bias_rel=bias, kmask=kmask, causal=causal, scale=scale, **cols)
    cuda_ops.attn_fwd(q, k, v, out=out, lse2=lse, **kw)
    lut = None
    if wb:
        lut = (torch.arange(Lq + Lk - 1) // 7).to(torch.int32).to(DEV)  # any monotone bucketisation exercises both paths
        if wb == "t5":
            lut = _t5ish_lut(Lq, Lk)
    res = []
    for ops in (cuda_ops, torch_ops):
        delta = torch.zeros(B, H, Lq, device=DEV)
        dq = torch.zeros(B * Lq, inner, device=DEV)
        dk = torch.zeros(B * Lk, 2 * inner, device=DEV, dtype=torch.bfloat16)
        dv = dk
        db = torch.zeros(H, Lq + Lk - 1, device=DEV) if wb else None
        ops.attn_bwd(q, k, v, out=out, lse2=lse, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dk, dk_col=0, dv=dv,
                     dv_col=inner, dbias_rel=db, bucket_lut=lut, **kw)
        res.append((dq, dk.clone(), db, delta))
    torch.cuda.synchronize()
    (dq, dkv, db, dl), (dq_r, dkv_r, db_r, dl_r) = res
    assert rel(dl, dl_r) < 1e-4
    assert rel(dq, dq_r) < 1.5e-2, rel(dq, dq_r)
    assert rel(dkv, dkv_r) < 1.5e-2, rel(dkv, dkv_r)
    if wb:
        if lut is not None:  # compare in bucket space (the uniform-tile fast path only preserves bucket sums)
            nb = int(lut.max().item()) + 1
            f = torch.zeros(nb, H, device=DEV).index_add_(0, lut.long(), db.t().contiguous())
            f_r = torch.zeros(nb, H, device=DEV).index_add_(0, lut.long(), db_r.t().contiguous())
            assert rel(f, f_r) < 1.5e-2, rel(f, f_r)


@pytest.mark.parametrize("M,N,K,mode", [(64, 2304, 768, "norm"), (64, 768, 768, "resid"), (64, 3072, 768, "norm_relu"),
                                        (64, 768, 3072, "resid"), (256, 1536, 1024, "norm"), (50, 1000, 768, "resid"),
                                        (7, 776, 4096, "plain")])
def test_decode_linear(cuda_ops, torch_ops, M, N, K, mode):
    """vc_decode_linear (decode2.cu): T5 RMS norm + nn.Linear (+ReLU / +residual) of one decode step in one launch."""
    g = gen(M + N + K)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV).bfloat16()
    x = torch.randn(M, K, generator=g).to(DEV)
    nw = (1 + 0.1 * torch.randn(K, generator=g)).to(DEV)
    if mode.startswith("norm"):
        out, ref = (torch.zeros(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(2))
        kw = dict(norm_w=nw, eps=1e-6, out_scale=0.5, relu=mode.endswith("relu"))
        cuda_ops.decode_linear(x, W, out, **kw)
        torch_ops.decode_linear(x, W, ref, **kw)
        assert rel(out, ref) < BF16_TOL
    elif mode == "resid":
        a = x.bfloat16()
        res = torch.randn(M, N, generator=g).to(DEV)
        out, ref = res.clone(), res.clone()
        cuda_ops.decode_linear(a, W, out, residual=out)          # in place, as the decode step uses it
        torch_ops.decode_linear(a, W, ref, residual=ref.clone())
        assert rel(out, ref) < F32_TOL
    else:
        a = x.bfloat16()
        out, ref = torch.zeros(M, N, device=DEV), torch.zeros(M, N, device=DEV)
        cuda_ops.decode_linear(a, W, out)
        torch_ops.decode_linear(a, W, ref)
        assert rel(out, ref) < F32_TOL


@pytest.mark.parametrize("kind", ["self", "cross", "cross_beams"])
def test_attn_single_query_decode(cuda_ops, torch_ops, kind):
    """Single-query attention over a KV cache (decode2.cu::attn_decode_kernel): causal self-attention at a device-side
    position with the relative bias row, cross-attention with a key mask, and beams sharing one copy of the K/V."""
    g = gen(31)
    H, inner = 12, 768
    if kind == "self":
        B, S, pos = 5, 256, 77
        q = (torch.randn(B, inner, generator=g) * 0.5).to(DEV).bfloat16()
        cache = (torch.randn(B * S, 2 * inner, generator=g) * 0.5).to(DEV).bfloat16()
        bias = torch.randn(H, 2 * S - 1, generator=g).to(DEV)
        pos_dev = torch.tensor([pos], dtype=torch.int32, device=DEV)
        kw = dict(q_col=0, k_col=0, v_col=inner, B=B, H=H, Lq=1, Lk=S, lse2=None, bias_rel=bias, kmask=None, causal=True,
                  scale=1.0, q_offset_dev=pos_dev, kv_batch_rows=S, bias_zero=S - 1, bias_len=2 * S - 1)
        out, ref = (torch.zeros(B, inner, device=DEV, dtype=torch.bfloat16) for _ in range(2))
        cuda_ops.attn_fwd(q, cache, cache, out=out, **kw)
        torch_ops.attn_fwd(q, cache, cache, out=ref, **kw)
    else:
        nb = 4 if kind == "cross_beams" else 1
        Bv, E = 3, 1100
        B = Bv * nb
        q = (torch.randn(B, inner, generator=g) * 0.5).to(DEV).bfloat16()
        kv = (torch.randn(Bv * E, 2 * inner, generator=g) * 0.5).to(DEV).bfloat16()
        lens = torch.tensor([1100, 640, 357])
        kmask = (torch.arange(E)[None, :] < lens[:, None]).to(torch.uint8).repeat_interleave(nb, 0).contiguous().to(DEV)
        kw = dict(q_col=0, k_col=0, v_col=inner, B=B, H=H, Lq=1, Lk=E, lse2=None, bias_rel=None, kmask=kmask, causal=False,
                  scale=1.0, kv_batch_div=nb if nb > 1 else 0)
        out, ref = (torch.zeros(B, inner, device=DEV, dtype=torch.bfloat16) for _ in range(2))
        cuda_ops.attn_fwd(q, kv, kv, out=out, **kw)
        torch_ops.attn_fwd(q, kv, kv, out=ref, **kw)
    assert rel(out, ref) < 8e-3, rel(out, ref)


@pytest.mark.parametrize("M,V,K,tn", [(4096, 32200, 768, 256), (300, 1012, 768, 128), (130, 1100, 256, 256)])
def test_fused_lm_head_cross_entropy(cuda_ops, M, V, K, tn):
    """LM head fused with F.cross_entropy(ignore_index=-100, label_smoothing=0.1) (modeling_t5.py:1714-1721): statistics
    GEMM (act 5) + vc_ce_combine + gradient GEMM (act 6) vs torch on the materialised logits; ragged vocabularies."""
    import torch.nn.functional as F
    g = gen(M + V)
    seq = (torch.randn(M, K, generator=g) * 0.4).to(DEV).bfloat16()
    W = (torch.randn(V, K, generator=g) * 0.4).to(DEV).bfloat16()
    labels = torch.randint(0, V, (M,), generator=g)
    labels[torch.rand(M, generator=g) < 0.2] = -100
    labels[0], labels[1] = V - 1, 0
    labels = labels.to(DEV)
    n_valid = (labels != -100).sum().float().reshape(1)
    stats = torch.empty(M, 2 * ((V + tn - 1) // tn), 3, device=DEV)
    zy, lse, loss = torch.empty(M, device=DEV), torch.empty(M, device=DEV), torch.empty(1, device=DEV)
    Vp = (V + 7) // 8 * 8
    dlogits = torch.zeros(M, Vp, device=DEV, dtype=torch.bfloat16)[:, :V]
    ce = dict(labels=labels, n_valid=n_valid, smoothing=0.1, stats=stats, zy=zy, lse=lse)
    cuda_ops.gemm(seq, W, None, act=5, tile_n=tn, ce=ce)
    cuda_ops.ce_combine(stats, zy, labels, n_valid, 0.1, V, lse, loss)
    cuda_ops.gemm(seq, W, dlogits, act=6, tile_n=tn, ce=ce)
    z = (seq.float() @ W.float().t()).requires_grad_(True)
    ref = F.cross_entropy(z, labels, ignore_index=-100, label_smoothing=0.1)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 2e-5 * abs(ref.item()), (loss.item(), ref.item())
    assert (lse - torch.logsumexp(z.detach(), -1)).abs().max().item() < 2e-4
    assert rel(dlogits, z.grad) < BF16_TOL, rel(dlogits, z.grad)
    assert float(dlogits[labels == -100].abs().sum()) == 0.0


@pytest.mark.parametrize("drop", [(0, 0), (0xBEEF, 6554)], ids=["nodrop", "drop0.1"])
def test_attn_skip_padded_query_tiles(cuda_ops, torch_ops, drop):
    """q_like_k (text-encoder self-attention): 128-query tiles past a sequence's last token are skipped.  Rows of real
    tokens must be unchanged, forward and backward; skipped rows read 0; with dO = 0 on the padding rows (what the train
    step guarantees: masked keys have exactly-zero probabilities downstream) dK/dV/dQ/d(bias) equal the full computation."""
    B, H, L = 3, 4, 1000
    g = gen(21)
    inner = H * 64
    qkv = (torch.randn(B * L, 3 * inner, generator=g) * 0.5).to(DEV).bfloat16()
    lens = torch.tensor([1000, 520, 130])
    kmask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.uint8).to(DEV)
    lut = _t5ish_lut(L, L)
    bias = torch.randn(33, H, generator=g).to(DEV)[lut.long()].t().contiguous()
    valid = kmask.bool().reshape(-1)
    kw = dict(q_col=0, k_col=inner, v_col=2 * inner, B=B, H=H, Lq=L, Lk=L, bias_rel=bias, kmask=kmask, causal=False,
              scale=1.0, drop=drop)
    out = torch.full((B * L, inner), 7.0, device=DEV, dtype=torch.bfloat16)
    ref = torch.zeros_like(out)
    lse, lse_ref = torch.zeros(B, H, L, device=DEV), torch.zeros(B, H, L, device=DEV)
    cuda_ops.attn_fwd(qkv, qkv, qkv, out=out, lse2=lse, q_like_k=True, **kw)
    torch_ops.attn_fwd(qkv, qkv, qkv, out=ref, lse2=lse_ref, **kw)
    assert rel(out[valid], ref[valid]) < 8e-3
    for b, n in enumerate(lens.tolist()):
        first_skipped = ((n - 1) // 128 + 1) * 128          # tiles entirely past the last token
        assert float(out.view(B, L, inner)[b, first_skipped:].abs().sum()) == 0.0
        assert (lse[b, :, :n] - lse_ref[b, :, :n]).abs().max().item() < 2e-3
    dout = (torch.randn(B * L, inner, generator=g) * 0.5).to(DEV).bfloat16()
    dout[~valid] = 0
    res = []
    for ops, o_, l_, extra in ((cuda_ops, out, lse, dict(q_like_k=True)), (torch_ops, ref, lse_ref, {})):
        delta = torch.zeros(B, H, L, device=DEV)
        dq = torch.zeros(B * L, inner, device=DEV)
        dkv = torch.zeros(B * L, 3 * inner, device=DEV, dtype=torch.bfloat16)
        db = torch.zeros(H, 2 * L - 1, device=DEV)
        ops.attn_bwd(qkv, qkv, qkv, out=o_, lse2=l_, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dkv, dk_col=inner,
                     dv=dkv, dv_col=2 * inner, dbias_rel=db, bucket_lut=lut, **kw, **extra)
        res.append((dq, dkv, db))
    (dq, dkv, db), (dq_r, dkv_r, db_r) = res
    assert rel(dq[valid], dq_r[valid]) < 1.5e-2 and float(dq[~valid].abs().sum()) == 0.0
    assert rel(dkv[valid][:, inner:], dkv_r[valid][:, inner:]) < 1.5e-2
    nb = int(lut.max().item()) + 1
    f = torch.zeros(nb, H, device=DEV).index_add_(0, lut.long(), db.t().contiguous())
    f_r = torch.zeros(nb, H, device=DEV).index_add_(0, lut.long(), db_r.t().contiguous())
    assert rel(f, f_r) < 1.5e-2


@pytest.mark.parametrize("kind,D,L", [(0, 768, 50), (1, 768, 50), (0, 1024, 50), (0, 256, 50), (0, 768, 3001), (1, 768, 2500),
                                      (1, 1024, 1777)])
def test_norms(cuda_ops, torch_ops, kind, D, L):
    g = gen(kind + D)
    B, T = 3, 10
    M = B * L
    x = torch.randn(M, D, generator=g).to(DEV)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
    b = (0.1 * torch.randn(D, generator=g)).to(DEV) if kind else None
    E = T + L
    outs = []
    for ops in (cuda_ops, torch_ops):
        ob = torch.zeros(B * E, D, device=DEV, dtype=torch.bfloat16)
        of = torch.zeros(B * E, D, device=DEV)
        rstd = torch.zeros(M, device=DEV)
        mean = torch.zeros(M, device=DEV)
        ops.norm_fwd(kind, x, w, b, out_bf16=ob, out_f32=of, rstd=rstd, mean=mean, eps=1e-5 if kind else 1e-6,
                     out_scale=0.5, rows_per_batch=L, out_batch_stride=E, out_row_offset=T)
        gy = torch.randn(B * E, D, generator=gen(7)).to(DEV)
        dx = torch.ones(M, D, device=DEV)
        dxb = torch.zeros(M, D, device=DEV, dtype=torch.bfloat16)
        dw = torch.zeros(D, device=DEV)
        db = torch.zeros(D, device=DEV)
        ops.norm_bwd(kind, gy if kind else gy.bfloat16(), x, w, rstd, mean, dx=dx, dx_bf16=dxb, accumulate_dx=True, dw=dw, db=db if kind else None,
                     scale=0.5, rows_per_batch=L, g_batch_stride=E, g_row_offset=T)
        outs.append((ob, of, rstd, mean, dx, dxb, dw, db))
    a, r = outs
    assert rel(a[1], r[1]) < 1e-5 and rel(a[0], r[0]) < BF16_TOL and rel(a[2], r[2]) < 1e-5
    assert rel(a[4], r[4]) < 1e-4 and rel(a[5], r[5]) < BF16_TOL and rel(a[6], r[6]) < 1e-4
    if kind:
        assert rel(a[7], r[7]) < 1e-4 and rel(a[3], r[3]) < 1e-4


def test_small_ops(cuda_ops, torch_ops):
    g = gen(3)
    V, d, n = 1100, 768, 300
    table = torch.randn(V, d, generator=g).to(DEV)
    ids = torch.randint(0, V, (n,), generator=g).to(DEV)
    B, S = 4, 33
    oids = torch.randint(0, 5, (B, S), generator=g).to(DEV)
    H, R = 12, 199
    lut = torch.randint(0, 32, (R,), generator=g).to(torch.int32).to(DEV)
    btab = torch.randn(32, H, generator=g).to(DEV)
    drel = torch.randn(H, R, generator=g).to(DEV)
    vid = torch.randn(B, 10, d, generator=g).to(DEV)
    pos = torch.randn(1, 100, d, generator=g).to(DEV)
    dvid = torch.randn(B * 10, d, generator=g).to(DEV)
    xb = torch.randn(300, 520, generator=g).to(DEV).bfloat16()
    res = []
    for ops in (cuda_ops, torch_ops):
        e = torch.zeros(n, d, device=DEV); ops.embed_fwd(ids, table, e)
        dt = torch.zeros(V, d, device=DEV); ops.embed_bwd(ids, e, dt)
        di = torch.zeros_like(oids); lb = torch.zeros_like(oids); nv = torch.zeros(1, device=DEV)
        ops.prepare_targets(oids, di, lb, nv)
        be = torch.zeros(H, R, device=DEV); ops.bias_expand(btab, lut, be)
        bf = torch.zeros(32, H, device=DEV); ops.bias_fold(drel, lut, bf)
        ap = torch.zeros_like(vid); ops.add_pos(vid, pos, ap, 100)
        dp = torch.zeros(1, 100, d, device=DEV); ops.add_pos_bwd(dvid, dp, B, 10, d, 100)
        cs = torch.zeros(520, device=DEV); ops.colsum_bf16(xb, cs)
        cb = torch.zeros(n, 2 * d, device=DEV, dtype=torch.bfloat16); ops.cast_f32_bf16(e, cb[:, d:], 0.5)
        mem = torch.zeros(B * 30, d, device=DEV, dtype=torch.bfloat16)
        ops.copy_rows_bf16(cb[:B * 10, :d].contiguous(), mem, B, 10, d, 30, 5)
        res.append((e, dt, di, lb, nv, be, bf, ap, dp, cs, cb, mem))
    for i, (a, r) in enumerate(zip(*res)):
        if a.dtype in (torch.int64,):
            assert torch.equal(a, r), i
        else:
            assert rel(a, r) < 1e-5 or (a - r).abs().max() < 1e-6, (i, rel(a, r))


@pytest.mark.parametrize("V", [32200, 32100, 1100])
def test_cross_entropy(cuda_ops, torch_ops, V):
    g = gen(9)
    n = 64
    Vp = (V + 7) // 8 * 8
    logits = torch.zeros(n, Vp, device=DEV)[:, :V]
    logits.copy_((torch.randn(n, V, generator=g) * 3).to(DEV))
    labels = torch.randint(0, V, (n,), generator=g).to(DEV)
    labels[::5] = -100
    nv = torch.tensor([float((labels != -100).sum())], device=DEV)
    out = []
    for ops in (cuda_ops, torch_ops):
        loss = torch.zeros(1, device=DEV)
        dl = torch.zeros(n, Vp, device=DEV, dtype=torch.bfloat16)[:, :V]
        ops.cross_entropy(logits, labels, nv, 0.1, loss, dl)
        out.append((loss, dl))
    ref = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-100, label_smoothing=0.1)
    assert abs(out[0][0].item() - ref.item()) < 1e-4 * abs(ref.item())
    assert abs(out[1][0].item() - ref.item()) < 1e-4 * abs(ref.item())
    assert rel(out[0][1], out[1][1]) < BF16_TOL


def test_optimizer_tail(cuda_ops, torch_ops):
    g = gen(4)
    n = 1100 * 768 + 4096
    p0 = torch.randn(n, generator=g).to(DEV)
    gr = (torch.randn(n, generator=g) * 0.01).to(DEV)
    res = []
    for ops in (cuda_ops, torch_ops):
        p = p0.clone(); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
        pb = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
        for step in (1, 2, 3):
            ns = torch.zeros(1, device=DEV)
            ops.sumsq(gr, ns)
            ops.adam_step(p, gr, m, v, pb, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, step=step, norm_sq=ns,
                          clip_max_norm=0.1, grad_scale=0.5)
        w = p[:1100 * 768].view(1100, 768)
        ops.renorm_time_tokens(w, pb[:1100 * 768].view(1100, 768), 100, torch.zeros(2, device=DEV))
        res.append((p, m, v, pb, ns))
    for a, r in zip(*res):
        assert rel(a, r) < 2e-5 if a.dtype == torch.float32 else rel(a, r) < BF16_TOL


# ----------------------------------------------------------------------------- dropout (counter-based masks)
DROP = (0xC0FFEE, 6554)   # p ~ 0.1


def test_dropout_gemm_epilogue(cuda_ops, torch_ops):
    M, N, K = 300, 2048, 768
    g = gen(21)
    A = (torch.randn(M, K, generator=g) * 0.3).to(DEV).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.3).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    resid = torch.randn(M, N, generator=g).to(DEV)
    aux = torch.randn(M, N, generator=g).to(DEV).bfloat16()
    for act in (0, 1, 2, 4):
        out = torch.empty(M, N, device=DEV)
        ref = torch.empty(M, N, device=DEV)
        kw = dict(bias=bias, residual=resid, act=act, aux=aux if act == 4 else None, drop=DROP)
        cuda_ops.gemm(A, B, out, **kw)
        torch_ops.gemm(A, B, ref, **kw)
        assert rel(out, ref) < 1e-4, act
        zeros = ((out - resid).abs() < 1e-12).float().mean().item()
        assert 0.05 < zeros < 0.6   # ~10% dropped (more with relu)


@pytest.mark.parametrize("preset", ["plain_bf16", "relu_bf16", "resid_f32", "relu_bwd_bf16", "gelu_bwd_bf16", "plain_f32"])
@pytest.mark.parametrize("drop", [False, True])
def test_gemm_epilogue_presets(cuda_ops, torch_ops, preset, drop):
    """The pair kernel's narrow epilogue instantiations (gemm_common.cuh epi_preset_mask): each call below needs only
    the features of one preset, so the host picks it instead of the all-features path the other epilogue tests take."""
    M, N, K = 1000, 1536, 768          # M >= 256: CTA-pair kernel; ragged last row block
    g = gen(31)
    A = (torch.randn(M, K, generator=g) * 0.3).to(DEV).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.3).to(DEV).bfloat16()
    resid = torch.randn(M, N, generator=g).to(DEV)
    aux = torch.randn(M, N, generator=g).to(DEV).bfloat16()
    bf = preset.endswith("bf16")
    kw = dict(drop=DROP) if drop else {}
    if preset == "relu_bf16":
        kw.update(act=1)
    elif preset == "resid_f32":
        kw.update(residual=resid, alpha=0.5)
    elif preset == "relu_bwd_bf16":
        kw.update(act=3, aux=aux)
    elif preset == "gelu_bwd_bf16":
        kw.update(act=4, aux=aux)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16 if bf else torch.float32)
    ref = torch.empty_like(out)
    cuda_ops.gemm(A, B, out, **kw)
    torch_ops.gemm(A, B, ref, **kw)
    assert rel(out, ref) < (BF16_TOL if bf else 1e-4)
    if drop and preset != "resid_f32":
        zeros = (out.float().abs() < 1e-12).float().mean().item()
        assert 0.05 < zeros < 0.65   # ~10 % dropped (about half more behind ReLU / its gradient)


@pytest.mark.parametrize("case", [ATTN_CASES[1], ATTN_CASES[2], ATTN_CASES[4], ATTN_CASES[7]])
def test_dropout_attention(cuda_ops, torch_ops, case):
    B, H, Lq, Lk, causal, wb, wm, scale, sa = case
    q, k, v, cols, bias, kmask = _attn_case(B, H, Lq, Lk, causal, wb, wm, scale, sa, seed=8)
    inner = H * 64
    kw = dict(B=B, H=H, Lq=Lq, Lk=Lk, bias_rel=bias, kmask=kmask, causal=causal, scale=scale, drop=DROP, **cols)
    outs = []
    dout = (torch.randn(B * Lq, inner, generator=gen(12)) * 0.5).to(DEV).bfloat16()
    for ops in (cuda_ops, torch_ops):
        out = torch.zeros(B * Lq, inner, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, H, Lq, device=DEV)
        ops.attn_fwd(q, k, v, out=out, lse2=lse, **kw)
        outs.append((out, lse))
    assert rel(outs[0][0], outs[1][0]) < 8e-3
    out, lse = outs[0]
    res = []
    for ops in (cuda_ops, torch_ops):
        delta = torch.zeros(B, H, Lq, device=DEV)
        dq = torch.zeros(B * Lq, inner, device=DEV)
        dk = torch.zeros(B * Lk, 2 * inner, device=DEV, dtype=torch.bfloat16)
        ops.attn_bwd(q, k, v, out=out, lse2=lse, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dk, dk_col=0, dv=dk,
                     dv_col=inner, dbias_rel=None, bucket_lut=None, **kw)
        res.append((dq, dk.clone()))
    assert rel(res[0][0], res[1][0]) < 1.5e-2 and rel(res[0][1], res[1][1]) < 1.5e-2


def test_dropout_norm_embed_pos(cuda_ops, torch_ops):
    g = gen(31)
    M, D = 200, 768
    x = torch.randn(M, D, generator=g).to(DEV)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
    gy = torch.randn(M, D, generator=g).to(DEV)
    table = torch.randn(500, D, generator=g).to(DEV)
    ids = torch.randint(0, 500, (M,), generator=g).to(DEV)
    vid = torch.randn(2, 100, D, generator=g).to(DEV)
    pos = torch.randn(1, 100, D, generator=g).to(DEV)
    res = []
    for ops in (cuda_ops, torch_ops):
        ob = torch.zeros(M, D, device=DEV, dtype=torch.bfloat16)
        rstd = torch.zeros(M, device=DEV)
        ops.norm_fwd(0, x, w, None, out_bf16=ob, rstd=rstd, eps=1e-6, drop=DROP)
        dx = torch.zeros(M, D, device=DEV)
        dxb = torch.zeros(M, D, device=DEV, dtype=torch.bfloat16)
        dw = torch.zeros(D, device=DEV)
        ops.norm_bwd(0, gy, x, w, rstd, None, dx=dx, dx_bf16=dxb, accumulate_dx=False, dw=dw, g_drop=DROP,
                     dxb_drop=(77, 13107))
        e = torch.zeros(M, D, device=DEV); ops.embed_fwd(ids, table, e, drop=DROP)
        dt = torch.zeros(500, D, device=DEV); ops.embed_bwd(ids, gy, dt, drop=DROP)
        ap = torch.zeros_like(vid); ops.add_pos(vid, pos, ap, 100, drop=DROP)
        dp = torch.zeros(1, 100, D, device=DEV); ops.add_pos_bwd(ap.view(200, D), dp, 2, 100, D, 100, drop=DROP)
        res.append((ob, dx, dxb, dw, e, dt, ap, dp))
    for i, (a, r) in enumerate(zip(*res)):
        assert rel(a, r) < (BF16_TOL if a.dtype == torch.bfloat16 else 1e-4), i
    assert 0.08 < (res[0][4] == 0).float().mean().item() < 0.12
    assert 0.17 < (res[0][2] == 0).float().mean().item() < 0.23   # dxb_drop p = 0.2


@pytest.mark.parametrize("B,nb,V", [(3, 4, 32200), (2, 1, 1100), (5, 8, 777)])
def test_beam_topk_and_kv_reorder(cuda_ops, torch_ops, B, nb, V):
    g = gen(41)
    Vp = (V + 7) // 8 * 8
    logits = (torch.randn(B * nb, Vp, generator=g) * 3).to(DEV)[:, :V]
    bs = (torch.randn(B * nb, generator=g) * 2).to(DEV)
    bs[1::2] = -1e9 if nb > 1 else bs[1::2]          # dead beams, as at the first step of HF beam search
    res = []
    for ops in (cuda_ops, torch_ops):
        s_ = torch.zeros(B, 2 * nb, device=DEV)
        t_ = torch.zeros(B, 2 * nb, device=DEV, dtype=torch.int32)
        b_ = torch.zeros(B, 2 * nb, device=DEV, dtype=torch.int32)
        ops.beam_topk(logits, bs, nb, s_, t_, b_)
        res.append((s_, t_, b_))
    torch.cuda.synchronize()
    (s0, t0, b0), (s1, t1, b1) = res
    live = s1 > -1e8                                   # candidates of dead beams tie at -1e9 (fp32): order unspecified
    assert torch.allclose(s0[live], s1[live], atol=2e-5, rtol=1e-6)
    assert torch.equal(t0[live], t1[live]) and torch.equal(b0[live], b1[live])
    assert bool((s0[~live] < -1e8).all())
    # cache reorder
    Bn, cap, C, n = B * nb, 24, 2 * 64 * 2, 17
    src = torch.randn(Bn, cap, C, generator=g).to(DEV).bfloat16()
    idx = torch.randint(0, Bn, (Bn,), generator=g).to(torch.int32).to(DEV)
    d0 = torch.zeros_like(src); d1 = torch.zeros_like(src)
    cuda_ops.kv_reorder(src, d0, idx, n)
    torch_ops.kv_reorder(src, d1, idx, n)
    torch.cuda.synchronize()
    assert torch.equal(d0, d1)
