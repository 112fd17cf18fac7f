"""CPU-side profile of one Vid2Seq.generate call at BASELINE configs[4] (where do the ~170 ms outside the decode loop go?)."""
import collections, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import DECODE_BATCH, DECODE_NEW, T_FRAMES, L_ASR, Tok, synth_batch
from vidchapters_b200 import Vid2Seq
dev = torch.device("cuda", 0)
m = Vid2Seq("t5-base", tokenizer=Tok(), seed=0, pretrained=False).to(dev).eval()
v, i, _ = [t.to(dev) for t in synth_batch(DECODE_BATCH, T_FRAMES, L_ASR, 8, 4321)]
tok = {"input_ids": i, "attention_mask": i != 0}
for _ in range(2):
    m.generate(v, tok, num_beams=1, max_length=DECODE_NEW)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter()
    m.generate(v, tok, num_beams=1, max_length=DECODE_NEW)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
print(f"generate wall {wall * 1e3:.1f} ms (CPU profiler on)")
os.makedirs("gpurun_out", exist_ok=True)
tr = "gpurun_out/gen.trace.json"; prof.export_chrome_trace(tr)
evs = json.load(open(tr))["traceEvents"]; os.remove(tr)
cpu = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    if "dur" in e and e.get("cat") in ("cpu_op", "cuda_runtime", "cuda_driver", "user_annotation"):
        cpu[e["name"][:60]][0] += 1; cpu[e["name"][:60]][1] += e["dur"]
for k, (n, d) in sorted(cpu.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{d / 1e3:9.2f} ms {n:5d}x  {k}")
