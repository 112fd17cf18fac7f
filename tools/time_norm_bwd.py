"""Time vc_norm_bwd at the config-2 shapes with an L2 flush between launches (CUDA events).  python tools/time_norm_bwd.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidchapters_b200.ops import CudaOps
ops = CudaOps()

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for kind, M, D in ((0, 16000, 768), (0, 4096, 768), (1, 1600, 768), (0, 8000, 1024)):
    x = torch.randn(M, D, device=dev); w = torch.ones(D, device=dev); g = torch.randn(M, D, device=dev).bfloat16()
    rstd = torch.rand(M, device=dev) + 0.5; mean = torch.zeros(M, device=dev)
    dx = torch.zeros(M, D, device=dev); dxb = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev) if kind else None
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.norm_bwd(kind, g, x, w, rstd, mean, dx=dx, dx_bf16=dxb, accumulate_dx=True, dw=dw, db=db)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    byts = M * D * (4 + 2 + 4 + 4 + 2)
    print(f"norm_bwd kind={kind} M={M} D={D}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  -> {byts / ts[len(ts)//2] / 1e6:.2f} TB/s algorithmic")
