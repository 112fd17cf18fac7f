"""Host input pipeline throughput: ours (vidchapters_b200.data) vs the reference's dataset code on the
same synthetic on-disk dataset, same stub tokenizer, one process, batches of 16.  The reference side runs only where
/root/reference exists.   python tools/bench_input_pipeline.py"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
from data_fixture import HFStubTokenizer, write_dataset
from vidchapters_b200 import data as D

torch.set_num_threads(1)
B, N = 16, 96
with tempfile.TemporaryDirectory() as d:
    js, feats, subs = write_dataset(d, n_videos=N, seed=5)
    tok = HFStubTokenizer()
    pad = {"input_tokens": 1000, "output_tokens": 256, "denoising_input_tokens": 1000, "denoising_output_tokens": 1000}
    rows = []

    def run(name, ds, collate):
        np.random.seed(0)
        t0 = time.time()
        ntok = 0
        for rep in range(3):
            for i in range(0, N, B):
                b = collate([ds[j] for j in range(i, i + B)])
                ntok += int((b["input_tokens"] != 0).sum()) + int((b["output_tokens"] != 0).sum()) + 100 * B
        dt = time.time() - t0
        rows.append((name, 3 * N / dt, ntok / dt, dt / (3 * N / B) * 1e3))

    ours = D.DenseVideoCaptioningDataset(js, feats, tokenizer=tok, subtitles_path=subs)
    run("ours (collate_dvc, longest)", ours, D.collate_dvc)
    run("ours (fixed shape 1000/256)", ours, lambda s: D.collate_dvc(s, pad_to=pad))
    pb = D.PinnedBatcher(B)
    run("ours (PinnedBatcher.fill)", ours, pb.fill)
    if os.path.isdir("/root/reference/dataset"):
        import importlib.util
        sys.path.insert(0, "/root/reference")
        spec = importlib.util.spec_from_file_location("ref_dvc_dataset", "/root/reference/dataset/dvc_dataset.py")
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        ref = mod.DenseVideoCaptioning_Dataset(js, feats, tokenizer=tok, subtitles_path=subs)
        run("reference (dvc_dataset.py)", ref, mod.densevideocaptioning_collate_fn)
print(f"# host input pipeline, 1 process / 1 thread, batches of {B}, {N} synthetic videos x 3 passes, stub tokenizer on both sides")
print(f"{'pipeline':32s} {'videos/s':>10s} {'tokens/s':>12s} {'ms/batch':>10s}")
for name, vps, tps, msb in rows:
    print(f"{name:32s} {vps:10.1f} {tps:12.0f} {msb:10.2f}")
