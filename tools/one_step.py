"""One eager train step of BASELINE configs[1] between cudaProfilerStart/Stop (driver for ncu launch lists):
   ncu --profile-from-start off --metrics ... python tools/one_step.py [--model t5-large] [--dropout 0.1]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import MODEL_BATCH, T_FRAMES, L_ASR, S_TGT, Tok, synth_batch
from vidchapters_b200 import Vid2Seq, Vid2SeqAdam

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="t5-base")
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--single-stream", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B = MODEL_BATCH[args.model]
m = Vid2Seq(args.model, tokenizer=Tok(), vis_drop=args.dropout, enc_drop=args.dropout, dec_drop=args.dropout, seed=0,
            pretrained=False).to(dev)
m.train()
opt = Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1, world_size=1)
if args.single_stream:
    m.engine.dual_stream = False
v, i, o = [t.to(dev) for t in synth_batch(B, T_FRAMES, L_ASR, S_TGT, 1234)]
tok = lambda x: {"input_ids": x, "attention_mask": x != 0}


def step():
    ld, _ = m(v, tok(i), tok(o))
    opt.zero_grad()
    ld["loss"].backward()
    opt.step()
    return ld["loss"]


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss.item(), "launches", m.engine.ops.launches)
