"""clock64 life of CTA (0,0,0) of the attention backward at a given shape (debug hook vc_debug_set_trace).
   python tools/attn_bwd_cta_life.py [cross|dec|enc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from vidchapters_b200.ops import CudaOps
from vidchapters_b200.engine import relative_position_bucket
ops = CudaOps(); dev = "cuda"; g = torch.Generator().manual_seed(0)
B, H = 16, 12; inner = H * 64
which = sys.argv[1] if len(sys.argv) > 1 else "cross"
Lq, Lk, causal, bias_on, mask_on, qlk = dict(cross=(256, 1100, False, False, True, False), dec=(256, 256, True, True, False, False),
                                             enc=(1000, 1000, False, True, True, True))[which]
q = (torch.randn(B * Lq, inner, generator=g) * 0.5).to(dev).bfloat16()
kv = (torch.randn(B * Lk, 2 * inner, generator=g) * 0.5).to(dev).bfloat16()
bias = lut = None
if bias_on:
    lut = relative_position_bucket(torch.arange(Lq + Lk - 1) - (Lq - 1), not causal).to(torch.int32).to(dev)
    bias = torch.randn(32, H, generator=g).to(dev)[lut.long()].t().contiguous()
kmask = None
if mask_on:
    lens = torch.randint(Lk // 2, Lk + 1, (B,), generator=g); lens[0] = Lk
    kmask = (torch.arange(Lk)[None] < lens[:, None]).to(torch.uint8).to(dev)
out = torch.zeros(B * Lq, inner, device=dev, dtype=torch.bfloat16); lse = torch.zeros(B, H, Lq, device=dev)
kw = dict(q_col=0, k_col=0, v_col=inner, B=B, H=H, Lq=Lq, Lk=Lk, bias_rel=bias, kmask=kmask, causal=causal, scale=1.0,
          drop=(0xC0FFEE, 6554), q_like_k=qlk)
dout = (torch.randn(B * Lq, inner, generator=g) * 0.5).to(dev).bfloat16()
delta = torch.zeros(B, H, Lq, device=dev); dq = torch.zeros(B * Lq, inner, device=dev)
dkv = torch.zeros(B * Lk, 2 * inner, device=dev, dtype=torch.bfloat16)
db = torch.zeros(H, Lq + Lk - 1, device=dev) if bias_on else None
ops.attn_fwd(q, kv, kv, out=out, lse2=lse, **kw)
bw = lambda: ops.attn_bwd(q, kv, kv, out=out, lse2=lse, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dkv, dk_col=0, dv=dkv,
                          dv_col=inner, dbias_rel=db, bucket_lut=lut, **kw)
bw(); bw()
trace = torch.zeros(2048, dtype=torch.int64, device=dev)
ops.lib.vc_debug_set_trace(C.c_void_p(trace.data_ptr()))
bw(); torch.cuda.synchronize()
ops.lib.vc_debug_set_trace(None)
names = {7: "tile_sum done", 8: "  dq: tmem ld ok", 9: "  dq: sts+fence", 10: "  dq: tma issued", 1: "iter top", 2: "s_full ok", 3: "math+pack done",
         4: "dq_full(i-1) ok", 5: "stores+arrive pds", 6: "dQ(i-1) staged", 100: "mma: wait pds", 101: "mma: pds ok",
         102: "mma: scores(i+1) issued", 103: "mma: dq_read ok", 104: "mma: dV/dK/dQ issued", 200: "CTA start", 201: "pdl_wait passed",
         202: "set-up done", 203: "mma: K/V landed", 204: "last dQ staged (reads done)", 205: "dV/dK stored", 206: "CTA end"}
ev = sorted((v & 0xFFFFFFFFFFFF, v >> 48, i) for i, v in enumerate(trace.cpu().tolist()) if v)
t0 = ev[0][0]; prev = t0
print(f"attention backward CTA (0,0,0), shape {which}: Lq={Lq} Lk={Lk}; cycles since CTA start (delta)")
for c, e, i in ev:
    print(f"  {names.get(e, e):30s} slot {i:5d} {c - t0:8d}  +{c - prev}")
    prev = c
