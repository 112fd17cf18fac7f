"""GPU busy time vs wall time of engine.encode at the decode batch (is the encoder phase of generate launch bound?)."""
import collections, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import DECODE_BATCH, T_FRAMES, L_ASR, Tok, synth_batch
from vidchapters_b200 import Vid2Seq
dev = torch.device("cuda", 0)
m = Vid2Seq("t5-base", tokenizer=Tok(), seed=0, pretrained=False).to(dev).eval()
v, i, _ = [t.to(dev) for t in synth_batch(DECODE_BATCH, T_FRAMES, L_ASR, 8, 4321)]
eng = m.engine; m._refresh_shadow()
with torch.no_grad():
    for _ in range(2):
        eng.encode(v, i, i != 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        eng.encode(v, i, i != 0)
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
os.makedirs("gpurun_out", exist_ok=True)
tr = "gpurun_out/enc.trace.json"; prof.export_chrome_trace(tr)
evs = json.load(open(tr))["traceEvents"]; os.remove(tr)
ev = sorted((e for e in evs if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e), key=lambda e: e["ts"])
busy = sum(e["dur"] for e in ev); span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print(f"wall (with profiler) {wall * 1e3:.1f} ms; GPU span {span / 1e3:.1f} ms; sum of kernel durations {busy / 1e3:.1f} ms; {len(ev)} GPU records")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e["name"].split("(")[0][-50:]][0] += 1; agg[e["name"].split("(")[0][-50:]][1] += e["dur"]
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{d / 1e3:8.2f} ms {n:4d}x  {k}")
cpu = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    if e.get("cat") in ("cpu_op", "cuda_runtime", "cuda_driver") and "dur" in e:
        cpu[e["name"][:50]][0] += 1; cpu[e["name"][:50]][1] += e["dur"]
print("CPU side:")
for k, (n, d) in sorted(cpu.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{d / 1e3:8.2f} ms {n:4d}x  {k}")
