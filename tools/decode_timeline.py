"""In-graph kernel timeline of the captured greedy decode step (CUPTI via torch.profiler; measurement tool only).
   python tools/decode_timeline.py"""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import DECODE_BATCH, DECODE_NEW, T_FRAMES, L_ASR, Tok, synth_batch
from vidchapters_b200 import Vid2Seq

dev = torch.device("cuda", 0)
m = Vid2Seq("t5-base", tokenizer=Tok(), seed=0, pretrained=False).to(dev).eval()
v, i, _ = [t.to(dev) for t in synth_batch(DECODE_BATCH, T_FRAMES, L_ASR, 8, 4321)]
eng = m.engine
m._refresh_shadow()
with torch.no_grad():
    memory, mem_mask, B, E = eng.encode(v, i, i != 0)
    st = eng._greedy_setup(memory, mem_mask, B, E, DECODE_NEW, True)
    for _ in range(100):
        st["graph"].replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            st["graph"].replay()
        torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
trace = "gpurun_out/decode_timeline.trace.json"
prof.export_chrome_trace(trace)
ev = sorted((e for e in json.load(open(trace))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e),
            key=lambda e: e["ts"])
os.remove(trace)
n = len(ev) // 3
step = ev[n:2 * n]
t0 = step[0]["ts"]; t1 = step[-1]["ts"] + step[-1]["dur"]
print(f"one decode step in the graph: {n} records, span {(t1 - t0):.1f} us")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
prev_end = t0
for e in step:
    a = e.get("args", {})
    k = (e["name"].split("(")[0][-40:], str(a.get("grid")))
    end = e["ts"] + e["dur"]
    agg[k][0] += 1; agg[k][1] += e["dur"]; agg[k][2] += max(0.0, end - prev_end); prev_end = max(prev_end, end)
print(f"{'sum dur':>9} {'end-to-end':>10} {'n':>4}  kernel grid")
for k, (c, d, ee) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{d:9.1f} {ee:10.1f} {c:4d}  {k[0]} {k[1]}  ({ee / c:.1f} us each)")
with open("gpurun_out/decode_timeline.kernels.txt", "w") as f:
    for e in step[: 40]:
        f.write(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} {e['name'][:60]} {e.get('args', {}).get('grid')}\n")
