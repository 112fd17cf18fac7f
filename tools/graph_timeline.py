"""In-graph kernel timeline of the graphed train step (torch.profiler / CUPTI activity records; measurement tool only).
Prints per-kernel-name totals of one replayed step, per-stream busy time and the idle gaps of the union timeline.
   python tools/graph_timeline.py [--model t5-base] [out_prefix]"""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import MODEL_BATCH, T_FRAMES, L_ASR, S_TGT, Tok, synth_batch
from vidchapters_b200 import GraphedTrainStep, Vid2Seq, Vid2SeqAdam

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="t5-base")
ap.add_argument("--out", default="gpurun_out/graph_timeline")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B = MODEL_BATCH[args.model]
m = Vid2Seq(args.model, tokenizer=Tok(), vis_drop=0.1, enc_drop=0.1, dec_drop=0.1, seed=0, pretrained=False).to(dev)
m.train()
opt = Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1, world_size=1)
v, i, o = [t.to(dev) for t in synth_batch(B, T_FRAMES, L_ASR, S_TGT, 1234)]
for _ in range(2):
    ld, _ = m(v, {"input_ids": i, "attention_mask": i != 0}, {"input_ids": o, "attention_mask": o != 0})
    opt.zero_grad(); ld["loss"].backward(); opt.step()
g = GraphedTrainStep(m, opt, v, i, o, warmup_steps=0)
for _ in range(5):
    g()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g()
    torch.cuda.synchronize()
trace = args.out + ".trace.json"
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
print(len(ev), "GPU activity records in 3 steps")
if not ev:
    sys.exit(0)
# split into steps by the largest two gaps between consecutive records of the adam kernel
adam = [e for e in ev if "adam_kernel" in e["name"]]
t_end = [e["ts"] + e["dur"] for e in adam]
lo, hi = t_end[0], t_end[1]          # the second step = (end of adam #1, end of adam #2]
step = [e for e in ev if lo < e["ts"] + e["dur"] <= hi + 1e-3]
t0 = min(e["ts"] for e in step); t1 = max(e["ts"] + e["dur"] for e in step)
print(f"step span {(t1 - t0) / 1e3:.3f} ms, {len(step)} records")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    n = e["name"].split("(")[0][-70:]
    agg[n][0] += 1; agg[n][1] += e["dur"]
tot = sum(a[1] for a in agg.values())
lines = [f"# in-graph kernel durations of ONE replayed train step ({args.model}); span {(t1 - t0) / 1e3:.3f} ms, sum of durations {tot / 1e3:.3f} ms"]
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{a[1]:10.1f} us {a[0]:5d}x {a[1] / a[0]:8.1f} us/launch {100 * a[1] / (t1 - t0):5.1f}% of span  {n}")
# union busy time and gaps
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in step)
busy = 0.0; cur_s, cur_e = iv[0]; gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; gaps.append((s - cur_e, cur_e - t0)); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
lines.append(f"# union busy {busy / 1e3:.3f} ms; idle {((t1 - t0) - busy) / 1e3:.3f} ms in {len(gaps)} gaps; largest: " +
             ", ".join(f"{g_:.1f}us@{at / 1e3:.2f}ms" for g_, at in sorted(gaps, reverse=True)[:8]))
streams = collections.defaultdict(float)
for e in step:
    streams[e.get("args", {}).get("stream", e.get("tid"))] += e["dur"]
lines.append("# per-stream sum of durations (ms): " + ", ".join(f"{k}: {v_ / 1e3:.2f}" for k, v_ in streams.items()))
# GEMM launches in time order with grid for a closer look
with open(args.out + ".kernels.txt", "w") as f:
    for e in step:
        a = e.get("args", {})
        f.write(f"{(e['ts'] - t0):10.1f} {e['dur']:8.1f} s{a.get('stream')} grid{a.get('grid')} {e['name'][:110]}\n")
open(args.out + ".txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
