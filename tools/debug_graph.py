import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vidchapters_b200 import GraphedTrainStep, Vid2Seq, Vid2SeqAdam
from vidchapters_b200 import lib as L

class Tok:
    pad_token_id, eos_token_id = 0, 1
    def __init__(self, n): self.n = n
    def __len__(self): return self.n

fx = torch.load("tests/golden/tiny.pt", weights_only=False)
cfg = fx["cfg"]
def build():
    m = Vid2Seq("t5-base", num_features=cfg["num_features"], depth=cfg["depth"], tokenizer=Tok(1100), dec_drop=0.0, t5_config=cfg)
    return m.to("cuda")
video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
it = {"input_ids": inp, "attention_mask": inp != 0}
ot = {"input_ids": out, "attention_mask": out != 0}
m2 = build(); m2.train()
o2 = Vid2SeqAdam(m2, lr=3e-4, clip_max_norm=0.1, world_size=1)
ld, _ = m2(video, it, ot); o2.zero_grad(); ld["loss"].backward(); o2.step()
torch.cuda.synchronize()
try:
    g = GraphedTrainStep(m2, o2, video, inp, out, warmup_steps=0)
    print("capture ok", g(video, inp, out).item())
except Exception as e:
    traceback.print_exc()
    print("LIB ERROR:", L.load().vc_last_error())
