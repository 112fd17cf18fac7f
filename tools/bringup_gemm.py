"""GPU bring-up for the tcgen05 GEMM: each variant runs in its own subprocess (a trap must not kill the sweep)."""
import json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [
    # name, M, N, K, a_mn, b_mn, tile_n, extra
    ("tn_small_64", 128, 64, 64, 0, 0, 64, {}),
    ("tn_small_128", 128, 128, 128, 0, 0, 128, {}),
    ("tn_small_256", 256, 256, 256, 0, 0, 256, {}),
    ("tn_ragged", 1600, 776, 768, 0, 0, 0, {}),
    ("tn_big", 16000, 3072, 768, 0, 0, 256, {}),
    ("nn_bmn_small", 128, 128, 128, 0, 1, 128, {}),
    ("nn_bmn_256", 256, 256, 256, 0, 1, 256, {}),
    ("nn_bmn_big", 4096, 768, 3072, 0, 1, 0, {}),
    ("tt_amn_small", 128, 128, 128, 1, 0, 128, {}),
    ("wgrad_small", 128, 128, 256, 1, 1, 128, {}),
    ("wgrad_big", 768, 3072, 16000, 1, 1, 0, {"splits": 8, "atomic": 1, "out_fp32": 1}),
    ("epi_bias_gelu", 1600, 2048, 768, 0, 0, 0, {"bias": 1, "act": 2}),
    ("epi_resid_fp32", 1600, 768, 2048, 0, 0, 0, {"bias": 1, "residual": 1, "out_fp32": 1}),
    ("epi_relu", 4096, 3072, 768, 0, 0, 0, {"act": 1}),
    ("epi_relu_bwd", 4096, 3072, 768, 0, 1, 0, {"act": 3}),
    ("lmhead", 4096, 32200, 768, 0, 0, 0, {"out_fp32": 1, "alpha": 0.5}),
    ("lmhead_dgrad", 4096, 768, 32200, 0, 1, 0, {}),
]


def run_one(name):
    import torch
    from vidchapters_b200 import ops as O
    spec = [v for v in VARIANTS if v[0] == name][0]
    _, M, N, K, a_mn, b_mn, tile_n, ex = spec
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).to(dev).bfloat16()
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    bias = torch.randn(N, generator=g).to(dev) if ex.get("bias") else None
    resid = torch.randn(M, N, generator=g).to(dev) if ex.get("residual") else None
    aux = torch.randn(M, N, generator=g).to(dev).bfloat16() if ex.get("act", 0) in (3, 4) else None
    alpha = ex.get("alpha", 1.0)
    ref = ref * alpha
    if bias is not None:
        ref = ref + bias
    act = ex.get("act", 0)
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    elif act == 3:
        ref = ref * (aux.float() > 0)
    if resid is not None:
        ref = ref + resid
    out_fp32 = ex.get("out_fp32", 0)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ops = O.CudaOps()
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bias, residual=resid, act=act, aux=aux, alpha=alpha,
              splits=ex.get("splits", 1), atomic=bool(ex.get("atomic", 0)), tile_n=tile_n)
    ops.gemm(A_st, B_st, out, **kw)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    # timing
    if ex.get("atomic"):
        ms = -1.0
    else:
        for _ in range(3):
            ops.gemm(A_st, B_st, out, **kw)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ops.gemm(A_st, B_st, out, **kw)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12 if ms > 0 else -1
    print(json.dumps(dict(name=name, max_abs=err, rel_l2=rel, ms=ms, tflops=tf)))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_one(sys.argv[1])
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    res = []
    for v in VARIANTS:
        t = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, v[0]], capture_output=True, text=True, timeout=120)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            r = json.loads(line[-1]) if line else dict(name=v[0], fail=p.returncode, err=(p.stderr or "")[-600:])
        except subprocess.TimeoutExpired:
            r = dict(name=v[0], fail="timeout")
        r["wall"] = round(time.time() - t, 1)
        print(r, flush=True)
        res.append(r)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bringup_gemm.json"), "w"), indent=1)
