"""GPU debug: run the engine with CudaOps and with TorchOps on the same device/weights; print where they diverge."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.torch_ops import TorchOps
from oracle import vid2seq_oracle as O
from vidchapters_b200.engine import Vid2SeqEngine
from vidchapters_b200.ops import CudaOps
from vidchapters_b200.init import init_state_dict

fx = torch.load("tests/golden/tiny.pt", weights_only=False)
cfg = fx["cfg"]
sd = init_state_dict(cfg, 0)
engs = []
for ops in (CudaOps(), TorchOps(flash_rounding=True)):
    e = Vid2SeqEngine(cfg, ops, "cuda")
    for n, t in sd.items():
        e.p(n).copy_(t.cuda())
    e.sync_bf16()
    engs.append(e)
video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
ctxs = [e.forward(video, inp, inp != 0, out, out != 0, want_logits=True)[1] for e in engs]
rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()
for i, (ra, rb) in enumerate(zip(ctxs[0]["tape"], ctxs[1]["tape"])):
    keys = [k for k in ("h", "qkv", "qc", "kv", "ctx", "lse", "act", "pre", "x0") if ra.get(k) is not None]
    print(i, ra["t"], " ".join(f"{k}={rel(ra[k], rb[k]):.2e}" for k in keys))
print("memory", rel(ctxs[0]["memory"], ctxs[1]["memory"]), "seq", rel(ctxs[0]["seq"], ctxs[1]["seq"]),
      "logits", rel(ctxs[0]["logits"], ctxs[1]["logits"]))
sdd = {k: v.cuda() for k, v in sd.items()}
o = O.vid2seq_forward(sdd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
print("cuda vs oracle-flash", rel(ctxs[0]["logits"].reshape(o["logits"].shape), o["logits"]),
      "torchops-flash vs oracle-flash", rel(ctxs[1]["logits"].reshape(o["logits"].shape), o["logits"]))
