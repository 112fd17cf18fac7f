"""Clock64 timeline of one CTA of the attention backward kernel (debug hook vc_debug_set_trace): where a query-tile
iteration spends its cycles — compute warp 2 (events 1..6) and the MMA thread (events 100..104)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from vidchapters_b200.ops import CudaOps
from vidchapters_b200.engine import relative_position_bucket

ops = CudaOps()
dev = "cuda"
g = torch.Generator().manual_seed(0)
B, H, L = 16, 12, 1000
inner = H * 64
qkv = (torch.randn(B * L, 3 * inner, generator=g) * 0.5).to(dev).bfloat16()
lut = relative_position_bucket(torch.arange(2 * L - 1) - (L - 1), True).to(torch.int32).to(dev)
bias = torch.randn(32, H, generator=g).to(dev)[lut.long()].t().contiguous()
lens = torch.randint(L // 2, L + 1, (B,), generator=g); lens[0] = L
kmask = (torch.arange(L)[None] < lens[:, None]).to(torch.uint8).to(dev)
out = torch.zeros(B * L, inner, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, L, device=dev)
drop = (0xC0FFEE, 6554) if os.environ.get("PROFILE_DROPOUT", "1") == "1" else (0, 0)
kw = dict(q_col=0, k_col=inner, v_col=2 * inner, B=B, H=H, Lq=L, Lk=L, bias_rel=bias, kmask=kmask, causal=False, scale=1.0,
          drop=drop, q_like_k=True)
dout = (torch.randn(B * L, inner, generator=g) * 0.5).to(dev).bfloat16()
delta = torch.zeros(B, H, L, device=dev); dq = torch.zeros(B * L, inner, device=dev)
dqkv = torch.zeros(B * L, 3 * inner, device=dev, dtype=torch.bfloat16)
db = torch.zeros(H, 2 * L - 1, device=dev)
trace = torch.zeros(2048, dtype=torch.int64, device=dev)
ops.attn_fwd(qkv, qkv, qkv, out=out, lse2=lse, **kw)
bw = lambda: ops.attn_bwd(qkv, qkv, qkv, out=out, lse2=lse, dout=dout, do_col=0, delta=delta, dq_acc=dq, dk=dqkv, dk_col=inner,
                          dv=dqkv, dv_col=2 * inner, dbias_rel=db, bucket_lut=lut, **kw)
bw(); bw()
if len(sys.argv) > 1 and sys.argv[1] == "fwd":
    ops.lib.vc_debug_set_trace(C.c_void_p(trace.data_ptr()))
    ops.attn_fwd(qkv, qkv, qkv, out=out, lse2=lse, **kw)
    torch.cuda.synchronize()
    ops.lib.vc_debug_set_trace(None)
    t = trace.cpu().tolist()
    ev = [(v >> 48, v & 0xFFFFFFFFFFFF, i) for i, v in enumerate(t) if v]
    t0 = min(e[1] for e in ev)
    nm = {1: "top", 2: "s_full ok", 3: "pass1 done", 4: "max exchanged", 5: "pv_done(j-1) ok (+rescale)", 6: "pass2 done, p_full",
          100: "mma: K ready", 101: "mma: s_free ok", 102: "mma: S(j+1) issued", 103: "mma: p_full ok", 104: "mma: PV issued"}
    for grp, lo, hi in (("softmax group A (warp 2)", 1, 19), ("softmax group B (warp 10)", 21, 39)):
        print(grp)
        prev = None
        for e, c, i in sorted([x for x in ev if lo <= x[0] <= hi], key=lambda x: x[1]):
            print(f"  tile {(i % 512) // 8}  {nm[e - (lo - 1)]:28s} {c - t0:8d}  +{(c - prev) if prev else 0}")
            prev = c
    print("MMA thread (A = group 0, B = group 1)")
    prev = None
    for e, c, i in sorted([x for x in ev if x[0] >= 100], key=lambda x: x[1]):
        gname = "B" if (e - 100) >= 10 else "A"
        print(f"  tile {(i - 1024) // 16} {gname}  {nm[100 + (e - 100) % 10]:24s} {c - t0:8d}  +{(c - prev) if prev else 0}")
        prev = c
    sys.exit(0)
ops.lib.vc_debug_set_trace(C.c_void_p(trace.data_ptr()))
bw()
torch.cuda.synchronize()
ops.lib.vc_debug_set_trace(None)
t = trace.cpu().tolist()
ev = [(v >> 48, v & 0xFFFFFFFFFFFF, i) for i, v in enumerate(t) if v and (v >> 48) < 200]   # (CTA-life events: attn_bwd_cta_life.py)
t0 = min(e[1] for e in ev)
names = {7: "tile_sum done", 8: "  dq: tmem ld ok", 9: "  dq: sts+fence", 10: "  dq: tma issued", 1: "iter top", 2: "s_full ok", 3: "math+pack done", 4: "dq_full(i-1) ok", 5: "stores+arrive pds", 6: "dQ(i-1) staged",
         100: "mma: wait pds", 101: "mma: pds ok", 102: "mma: scores(i+1) issued", 103: "mma: dq_read ok", 104: "mma: dV/dK/dQ issued"}
print("compute warp 2 (cycles since first event; delta to previous):")
prev = None
for e, c, i in sorted([x for x in ev if x[0] < 100], key=lambda x: x[1]):
    print(f"  iter {i // 8 if i < 512 else (i - 512) // 4 + 1}  {names[e]:22s} {c - t0:8d}  +{(c - prev) if prev else 0}")
    prev = c
print("MMA thread:")
prev = None
for e, c, i in sorted([x for x in ev if x[0] >= 100], key=lambda x: x[1]):
    print(f"  iter {(i - 1024) // 8}  {names[e]:26s} {c - t0:8d}  +{(c - prev) if prev else 0}")
    prev = c
