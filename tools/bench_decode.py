"""Decode throughput (SURVEY §8d secondary row, cfg5: t5-base, batch 64, 100 frames + 1000 ASR tokens of memory, 256 new
tokens): greedy and beam-4 `generate` through the drop-in module.  Random-init weights never emit eos, so every sequence
runs the full 256 steps.  Prints one JSON line per mode.   python tools/bench_decode.py [--batch 64] [--steps 256]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vidchapters_b200 import T5_BASE, Vid2Seq


class Tok:
    pad_token_id, eos_token_id = 0, 1

    def __init__(self, n): self.n = n
    def __len__(self): return self.n
    def batch_decode(self, ids, skip_special_tokens=True): return ["" for _ in ids]


ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=256)
args = ap.parse_args()
cfg = dict(T5_BASE)
m = Vid2Seq("t5-base", num_features=100, tokenizer=Tok(cfg["base_vocab"] + cfg["num_bins"]), t5_config=cfg, seed=0).to("cuda").eval()
g = torch.Generator().manual_seed(5)
B, T, L = args.batch, 100, 1000
video = torch.randn(B, T, 768, generator=g).cuda()
inp = torch.randint(2, 32100, (B, L), generator=g)
for b in range(B):
    n = int(torch.randint(L // 2, L + 1, (1,), generator=g)); inp[b, n:] = 0
inp = inp.cuda()
it = {"input_ids": inp, "attention_mask": inp != 0}
for nb in (1, 4):
    m.generate(video, it, num_beams=nb, max_length=8)        # warm-up (allocations, graph capture path)
    torch.cuda.synchronize()
    t0 = time.time()
    m.generate(video, it, num_beams=nb, max_length=args.steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    n_tok = m.last_generated_ids.shape[1] - 1
    print(json.dumps({"metric": "decode_tokens_per_sec", "mode": "greedy" if nb == 1 else f"beam{nb}", "value": B * n_tok / dt,
                      "unit": "tokens/s", "batch": B, "new_tokens": n_tok, "seconds": dt,
                      "includes": "visual + text encoder, cross K/V projection, graph capture, decode loop"}), flush=True)
