"""Synthetic on-disk dataset in the reference's formats (annotation json, per-video .npy features, ASR
pickle) + an HF-call-compatible stub tokenizer.  Shared by oracle/make_golden_data.py (mints the golden outputs from the
reference's own dataset code) and tests/test_data_pipeline_cpu.py."""
import json
import os
import pickle
import zlib

import numpy as np
import torch

WORDS = ("add the onions to pan and stir until golden then pour water over rice cover with a lid wait ten minutes "
         "chapter intro outro thanks for watching subscribe next we cut slice mix bake serve").split()


class HFStubTokenizer:
    """Word-hash tokenizer with the call signature dvc_dataset.py uses (T5Tokenizer is third-party, no spiece.model
    offline): ids in [2, base_vocab); len() counts the 100 time tokens like the reference's extended tokenizer."""
    pad_token_id, eos_token_id = 0, 1

    def __init__(self, base_vocab=32100, num_bins=100):
        self.base, self.n = base_vocab, base_vocab + num_bins

    def __len__(self):
        return self.n

    def __call__(self, text, add_special_tokens=False, max_length=None, padding="do_not_pad", truncation=True,
                 return_tensors="pt"):
        ids = [2 + zlib.crc32(w.encode()) % (self.base - 2) for w in text.replace(".", " .").split()]
        if truncation and max_length is not None:
            ids = ids[:max_length]
        return {"input_ids": torch.tensor([ids], dtype=torch.long)}


def _sentence(rng, lo, hi):
    return " ".join(rng.choice(WORDS) for _ in range(int(rng.integers(lo, hi))))


def write_dataset(root, n_videos=12, seed=0, features_dim=768):
    """Writes root/{ann.json, feats/<id>.npy, subs.pkl}; covers: more / fewer / exactly max_feats frames, videos without
    ASR, ASR outside [0, duration], long ASR (truncation at max_input_tokens), many chapters (truncation at
    max_output_tokens)."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "feats"), exist_ok=True)
    ann, subs = {}, {}
    for v in range(n_videos):
        vid = f"vid{v:08d}"                      # 11 characters, like a YouTube id
        duration = float(rng.integers(60, 1800))
        n_frames = [100, 37, 250, 1, 100][v % 5] if v < 5 else int(rng.integers(5, 400))
        np.save(os.path.join(root, "feats", vid + ".npy"), rng.standard_normal((n_frames, features_dim)).astype(np.float32))
        n_ch = 40 if v == 3 else int(rng.integers(1, 8))
        cuts = np.sort(rng.uniform(0, duration, size=n_ch + 1))
        ann[vid] = {"duration": duration, "timestamps": [[float(cuts[i]), float(cuts[i + 1])] for i in range(n_ch)],
                    "sentences": [_sentence(rng, 2, 9) for _ in range(n_ch)]}
        if v % 4 == 1:
            continue                             # no ASR for this video
        n_sub = 300 if v == 2 else int(rng.integers(1, 40))
        st = np.sort(rng.uniform(-5 if v == 6 else 0, duration, size=n_sub))
        ed = st + rng.uniform(1, 8, size=n_sub)
        if v == 7:                               # every subtitle outside the video: treated as "no subtitles"
            st, ed = st + 2 * duration, ed + 2 * duration
        subs[vid] = {"start": [float(x) for x in st], "end": [float(x) for x in ed],
                     "text": [_sentence(rng, 1, 12) for _ in range(n_sub)]}
    with open(os.path.join(root, "ann.json"), "w") as f:
        json.dump(ann, f)
    with open(os.path.join(root, "subs.pkl"), "wb") as f:
        pickle.dump(subs, f)
    return os.path.join(root, "ann.json"), os.path.join(root, "feats"), os.path.join(root, "subs.pkl")


def write_yt_dataset(root, n_videos=10, seed=1, features_dim=768):
    """Pretraining-set layout of dataset/yt_dataset.py: root/{list.csv, feats/<path>.npy, subs/<id>.pkl}; covers subtitles
    with and without a stored duration, out-of-range subtitles, a video whose subtitles are all dropped, long transcripts."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "feats"), exist_ok=True)
    os.makedirs(os.path.join(root, "subs"), exist_ok=True)
    rows = []
    for v in range(n_videos):
        vid = f"ytv{v:08d}"
        n_frames = [100, 60, 333, 7][v % 4] if v < 4 else int(rng.integers(5, 400))
        np.save(os.path.join(root, "feats", vid + ".npy"), rng.standard_normal((n_frames, features_dim)).astype(np.float32))
        n_sub = 260 if v == 1 else int(rng.integers(1, 50))
        horizon = float(n_frames + 1)
        st = np.sort(rng.uniform(-3 if v == 2 else 0, horizon, size=n_sub))
        ed = st + rng.uniform(0.5, 6, size=n_sub)
        if v == 3:
            st, ed = st + 10 * horizon, ed + 10 * horizon       # nothing survives the duration filter
        sub = {"start": [float(x) for x in st], "end": [float(x) for x in ed], "text": [_sentence(rng, 1, 12) for _ in range(n_sub)]}
        if v % 3 == 0:
            sub["duration"] = float(n_frames) * 1.5
        with open(os.path.join(root, "subs", vid + ".pkl"), "wb") as f:
            pickle.dump(sub, f)
        rows.append({"video_id": vid, "video_path": vid + ".npy"})
    pd.DataFrame(rows).to_csv(os.path.join(root, "list.csv"), index=False)
    return os.path.join(root, "list.csv"), os.path.join(root, "feats"), os.path.join(root, "subs")
