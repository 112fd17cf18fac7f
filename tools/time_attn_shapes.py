"""Forward / backward attention time at the three config-2 shapes (encoder self, decoder self, decoder cross), CUDA events,
20 back-to-back launches each.   python tools/time_attn_shapes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vidchapters_b200.ops import CudaOps
from vidchapters_b200.engine import relative_position_bucket

ops = CudaOps()
dev = "cuda"
g = torch.Generator().manual_seed(0)
B, H = 16, 12
inner = H * 64
drop = (0xC0FFEE, 6554)


def run(name, Lq, Lk, causal, bias_on, mask_on, qlk):
    q = (torch.randn(B * Lq, inner, generator=g) * 0.5).to(dev).bfloat16()
    kv = (torch.randn(B * Lk, 2 * inner, generator=g) * 0.5).to(dev).bfloat16()
    bias = lut = None
    if bias_on:
        lut = relative_position_bucket(torch.arange(Lq + Lk - 1) - (Lq - 1), not causal).to(torch.int32).to(dev)
        bias = torch.randn(32, H, generator=g).to(dev)[lut.long()].t().contiguous()
    kmask = None
    if mask_on:
        lens = torch.randint(Lk // 2, Lk + 1, (B,), generator=g)
        kmask = (torch.arange(Lk)[None] < lens[:, None]).to(torch.uint8).to(dev)
    out = torch.zeros(B * Lq, inner, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Lq, device=dev)
    kw = dict(q_col=0, k_col=0, v_col=inner, B=B, H=H, Lq=Lq, Lk=Lk, bias_rel=bias, kmask=kmask, causal=causal, scale=1.0,
              drop=drop, q_like_k=qlk)
    dout = (torch.randn(B * Lq, inner, generator=g) * 0.5).to(dev).bfloat16()
    delta = torch.zeros(B, H, Lq, device=dev); dq = torch.zeros(B * Lq, inner, device=dev)
    dkv = torch.zeros(B * Lk, 2 * inner, device=dev, dtype=torch.bfloat16)
    db = torch.zeros(H, Lq + Lk - 1, device=dev) if bias_on else None
    fw = lambda: ops.attn_fwd(q, kv, kv, out=out, lse2=lse, **kw)
    bw = lambda: ops.attn_bwd(q, kv, kv, out=out, lse2=lse, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dkv, dk_col=0,
                              dv=dkv, dv_col=inner, dbias_rel=db, bucket_lut=lut, **kw)
    for nm, fn in (("fwd", fw), ("bwd(+delta)", bw)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:14s} Lq={Lq:5d} Lk={Lk:5d} {nm:12s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us")


run("encoder self", 1000, 1000, False, True, True, True)
run("decoder self", 256, 256, True, True, False, False)
run("decoder cross", 256, 1100, False, False, True, False)
