"""One eager greedy decode step of BASELINE configs[4] (batch 64, 1100 memory tokens) between cudaProfilerStart/Stop:
   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum python tools/one_decode_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import DECODE_BATCH, DECODE_NEW, T_FRAMES, L_ASR, Tok, synth_batch
from vidchapters_b200 import Vid2Seq

dev = torch.device("cuda", 0)
m = Vid2Seq("t5-base", tokenizer=Tok(), seed=0, pretrained=False).to(dev).eval()
v, i, _ = [t.to(dev) for t in synth_batch(DECODE_BATCH, T_FRAMES, L_ASR, 8, 4321)]
eng = m.engine
m._refresh_shadow()
with torch.no_grad():
    memory, mem_mask, B, E = eng.encode(v, i, i != 0)
    st = eng._greedy_setup(memory, mem_mask, B, E, DECODE_NEW, use_graph=False)
    for _ in range(100):          # position 100: a representative self-attention cache length
        st["step"]()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    st["step"]()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
