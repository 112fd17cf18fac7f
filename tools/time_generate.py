"""Wall-clock split of Vid2Seq.generate at BASELINE configs[4]: encoders / greedy set-up (cross K/V, warm-up step, graph
capture) / decode loop.   python tools/time_generate.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import DECODE_BATCH, DECODE_NEW, T_FRAMES, L_ASR, Tok, synth_batch
from vidchapters_b200 import Vid2Seq

dev = torch.device("cuda", 0)
m = Vid2Seq("t5-base", tokenizer=Tok(), seed=0, pretrained=False).to(dev).eval()
v, i, _ = [t.to(dev) for t in synth_batch(DECODE_BATCH, T_FRAMES, L_ASR, 8, 4321)]
eng = m.engine
m._refresh_shadow()
sync = lambda: torch.cuda.synchronize()
with torch.no_grad():
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        memory, mem_mask, B, E = eng.encode(v, i, i != 0)
        sync(); t1 = time.perf_counter()
        st = eng._greedy_setup(memory, mem_mask, B, E, DECODE_NEW, True)
        sync(); t2 = time.perf_counter()
        for n in range(DECODE_NEW):
            if n and n % 16 == 0 and bool(st["done"].all().item()):
                break
            st["graph"].replay()
        sync(); t3 = time.perf_counter()
        print(f"rep {rep}: encode {1e3 * (t1 - t0):.1f} ms, set-up+capture {1e3 * (t2 - t1):.1f} ms, loop {1e3 * (t3 - t2):.1f} ms ({n + 1} steps)")
    sync(); t0 = time.perf_counter()
    m.generate(v, {"input_ids": i, "attention_mask": i != 0}, num_beams=1, max_length=DECODE_NEW)
    sync(); print(f"Vid2Seq.generate: {1e3 * (time.perf_counter() - t0):.1f} ms")
