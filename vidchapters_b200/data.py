"""Host input pipeline of the dvc.py train step: per-video CLIP features + ASR + chapter annotations
-> the reference's sample dict -> padded (optionally fixed-shape, pinned) batches for `Vid2Seq.forward`.

Replaces, with the same inputs, outputs and random-number consumption:
  dataset/dvc_dataset.py:61-84    _get_video          -> subsample_pad_features
  dataset/dvc_dataset.py:86-89    time_tokenize       -> time_tokenize
  dataset/dvc_dataset.py:91-165   __getitem__         -> DenseVideoCaptioningDataset.__getitem__
  dataset/dvc_dataset.py:168-208  collate             -> collate_dvc / PinnedBatcher
  dataset/yt_dataset.py:10-165    pretraining dataset -> YTDataset (+ collate_dvc, which also covers yt_collate_fn)
  util/t5.py:3-94                 span corruption     -> random_spans_noise_mask / span_corrupt
Integer work throughout: parity with the reference is bit-exact (tests/test_data_pipeline_cpu.py, golden minted from
the reference by oracle/make_golden_data.py).  The tokenizer (sentencepiece T5Tokenizer, third-party) stays an injected
callable with the HF call signature the reference uses.

Why it is here: at ~0.7 M tokens/s per GPU a 16-video batch is consumed every 32 ms; the reference's per-sample torch.cat
chains and Python loops (one small tensor per timestamp, per caption, per padded row) cost more than that per batch.
This version keeps every sample as flat numpy int64 until the batch is written once into preallocated pinned buffers.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


# ----------------------------------------------------------------------------- features
def subsample_pad_features(video, max_feats: int, features_dim: int) -> torch.Tensor:
    """[N, D] features -> [max_feats, D] float32: uniform temporal subsampling video[(j*N)//max_feats] when N > max_feats,
    zero padding when N < max_feats (dvc_dataset.py:61-84)."""
    v = torch.as_tensor(video).float()
    n = v.shape[0]
    if n > max_feats:
        idx = (torch.arange(max_feats, dtype=torch.int64) * n) // max_feats
        return v.index_select(0, idx)
    if n < max_feats:
        out = torch.zeros(max_feats, features_dim, dtype=torch.float32)
        out[:n] = v
        return out
    return v


def time_tokenize(x: float, duration: float, num_bins: int, num_text_tokens: int) -> int:
    """Quantised time token id (dvc_dataset.py:86-89): int((num_bins-1)*x / duration) + num_text_tokens."""
    t = int(float((num_bins - 1) * x) / float(duration))
    assert t <= num_bins
    return t + num_text_tokens


def _clean_text(text: str) -> str:  # dvc_dataset.py:54-59
    text = text.strip().capitalize()
    return text if text[-1] == "." else text + "."


def _tokenize(tokenizer, text: str, max_length: int) -> np.ndarray:
    ids = tokenizer(text, add_special_tokens=False, max_length=max_length, padding="do_not_pad", truncation=True,
                    return_tensors="pt")["input_ids"][0]
    return np.asarray(ids, dtype=np.int64)


def timed_token_sequence(starts: Sequence[float], ends: Sequence[float], texts: Sequence[str], duration: float, tokenizer,
                         num_bins: int, num_text_tokens: int, max_tokens: int) -> np.ndarray:
    """[t_start, t_end, text tokens ...] per segment, concatenated, cut to max_tokens - 1, + eos
    (dvc_dataset.py:110-122 for ASR, :146-158 for chapters)."""
    parts: List[np.ndarray] = []
    for st, ed, tx in zip(starts, ends, texts):
        parts.append(np.array([time_tokenize(st, duration, num_bins, num_text_tokens),
                               time_tokenize(ed, duration, num_bins, num_text_tokens)], dtype=np.int64))
        parts.append(_tokenize(tokenizer, tx, max_tokens))
    seq = np.concatenate(parts)[:max_tokens - 1]
    return np.concatenate([seq, np.array([tokenizer.eos_token_id], dtype=np.int64)])


# ----------------------------------------------------------------------------- span corruption (util/t5.py)
def _random_segmentation(num_items: int, num_segments: int, rng) -> np.ndarray:
    """Lengths of a uniformly random partition of num_items into num_segments non-empty runs.  Consumes the random
    stream exactly like util/t5.py:60-74 (one shuffle of a boolean vector of length num_items - 1)."""
    cuts = np.arange(num_items - 1) < (num_segments - 1)
    rng.shuffle(cuts)
    bounds = np.flatnonzero(cuts) + 1                       # positions where a new segment starts
    return np.diff(np.concatenate([[0], bounds, [num_items]]))


def random_spans_noise_mask(length: int, noise_density: float, mean_noise_span_length: float, rng=np.random) -> np.ndarray:
    """Boolean [length] mask of noise tokens: alternating non-noise / noise spans starting with non-noise
    (util/t5.py:36-94, T5's random_spans_helper)."""
    num_noise = int(np.round(length * noise_density))
    num_noise = min(max(num_noise, 1), length - 1)
    num_spans = max(int(np.round(num_noise / mean_noise_span_length)), 1)
    noise_len = _random_segmentation(num_noise, num_spans, rng)
    nonnoise_len = _random_segmentation(length - num_noise, num_spans, rng)
    lengths = np.stack([nonnoise_len, noise_len], axis=1).reshape(-1)      # non-noise, noise, non-noise, ...
    return np.repeat(np.arange(2 * num_spans) % 2 == 1, lengths)[:length]


def _replace_spans(tokens: np.ndarray, drop: np.ndarray, first_sentinel: int, eos: int) -> np.ndarray:
    """Every maximal run of `drop` positions collapses into one sentinel (first_sentinel, first_sentinel-1, ...), the other
    tokens are kept, eos is appended (util/t5.py:3-33: create_sentinel_ids + filter_input_ids)."""
    starts = drop & ~np.concatenate([[False], drop[:-1]])
    sentinels = first_sentinel - (np.cumsum(starts) - 1)
    keep = ~drop | starts
    out = np.where(starts, sentinels, tokens)[keep]
    return np.concatenate([out, np.array([eos], dtype=np.int64)]).astype(np.int64)


def span_corrupt(tokens: np.ndarray, vocab_len: int, num_bins: int, eos: int, noise_density: float = 0.25,
                 mean_noise_span_length: float = 5, rng=np.random):
    """(denoising_input, denoising_output) of T5 span corruption over `tokens` (dvc_dataset.py:126-139); the k-th
    sentinel is vocab_len - num_bins - k, i.e. <extra_id_{k-1}> below the time tokens."""
    noise = random_spans_noise_mask(len(tokens), noise_density, mean_noise_span_length, rng)
    first = vocab_len - num_bins - 1
    return _replace_spans(tokens, noise, first, eos), _replace_spans(tokens, ~noise, first, eos)


# ----------------------------------------------------------------------------- dataset
class DenseVideoCaptioningDataset(torch.utils.data.Dataset):
    """Drop-in for dataset/dvc_dataset.py::DenseVideoCaptioning_Dataset (same constructor, same sample dict)."""

    def __init__(self, json_path, features_path, max_feats=100, features_dim=768, tokenizer=None, subtitles_path=None,
                 num_bins=100, max_input_tokens=1000, max_output_tokens=256, noise_density=0.25, mean_noise_span_length=5):
        with open(json_path, "r") as f:
            self.data = json.load(f)
        self.vids = list(self.data.keys())
        self.features, self.features_path = None, None
        if os.path.isdir(features_path):
            self.features_path = features_path
        else:
            self.features = torch.load(features_path)
        self.subs, self.subs_path = None, None
        if subtitles_path and os.path.isdir(subtitles_path):
            self.subs_path = subtitles_path
        elif subtitles_path and os.path.exists(subtitles_path):
            with open(subtitles_path, "rb") as f:
                self.subs = pickle.load(f)
        self.max_feats, self.features_dim, self.tokenizer = max_feats, features_dim, tokenizer
        self.num_bins = num_bins
        self.max_input_tokens, self.max_output_tokens = max_input_tokens, max_output_tokens
        self.num_text_tokens = len(tokenizer) - num_bins
        self.noise_density, self.mean_noise_span_length = noise_density, mean_noise_span_length

    def __len__(self):
        return len(self.data)

    def _features(self, vid: str):
        if self.features is not None:
            return self.features[vid]
        p = os.path.join(self.features_path, vid + ".mp4.npy")
        if not os.path.exists(p):
            p = os.path.join(self.features_path, vid + ".npy")
        return np.load(p)

    def _subtitles(self, video_id: str):
        key = video_id[-11:]
        if self.subs is not None and key in self.subs:
            return self.subs[key]
        if self.subs_path is not None and os.path.exists(os.path.join(self.subs_path, video_id + ".pkl")):
            with open(os.path.join(self.subs_path, key + ".pkl"), "rb") as f:
                return pickle.load(f)
        return None

    def __getitem__(self, idx):
        video_id = self.vids[idx]
        ann = self.data[video_id]
        duration = ann["duration"]
        tok, eos = self.tokenizer, self.tokenizer.eos_token_id
        video = subsample_pad_features(self._features(video_id[-11:]), self.max_feats, self.features_dim)
        inp = np.array([eos], dtype=np.int64)
        sub = self._subtitles(video_id)
        if sub is not None:
            keep = [i for i, (x, y) in enumerate(zip(sub["start"], sub["end"])) if x >= 0 and y <= duration]
            if keep:
                inp = timed_token_sequence([sub["start"][i] for i in keep], [sub["end"][i] for i in keep],
                                           [_clean_text(sub["text"][i]) for i in keep], duration, tok, self.num_bins,
                                           self.num_text_tokens, self.max_input_tokens)
        if len(inp) > 1:
            den_in, den_out = span_corrupt(inp, len(tok), self.num_bins, eos, self.noise_density, self.mean_noise_span_length)
        else:
            den_in, den_out = np.array([0], dtype=np.int64), inp
        out = timed_token_sequence([t[0] for t in ann["timestamps"]], [t[1] for t in ann["timestamps"]],
                                   [_clean_text(x) for x in ann["sentences"]], duration, tok, self.num_bins,
                                   self.num_text_tokens, self.max_output_tokens)
        f = torch.from_numpy
        return {"video_id": video_id, "duration": duration, "video": video, "input_tokens": f(inp), "output_tokens": f(out),
                "denoising_input_tokens": f(den_in), "denoising_output_tokens": f(den_out)}


class YTDataset(torch.utils.data.Dataset):
    """Drop-in for dataset/yt_dataset.py::YT_Dataset, the pretraining set (HowTo100M / VidChapters ASR): the "output" of
    the generative pass is the timed transcript itself, plus its span-corrupted pair (yt_dataset.py:84-131)."""

    def __init__(self, csv_path, features_path, subtitles_path, max_feats=100, features_dim=768, tokenizer=None, num_bins=100,
                 max_input_tokens=1000, max_output_tokens=1000, noise_density=0.25, mean_noise_span_length=5):
        import pandas as pd
        self.data = pd.read_csv(csv_path)
        self.features_path, self.subtitles_path = features_path, subtitles_path
        self.max_feats, self.features_dim, self.tokenizer = max_feats, features_dim, tokenizer
        self.num_bins = num_bins
        self.max_input_tokens, self.max_output_tokens = max_input_tokens, max_output_tokens
        self.num_text_tokens = len(tokenizer) - num_bins
        self.noise_density, self.mean_noise_span_length = noise_density, mean_noise_span_length

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        video_id = self.data["video_id"][idx]
        with open(os.path.join(self.subtitles_path, video_id + ".pkl"), "rb") as f:
            sub = pickle.load(f)
        raw = np.load(os.path.join(self.features_path, self.data["video_path"][idx]))
        duration = sub["duration"] if "duration" in sub else len(raw) + 1          # yt_dataset.py:53-55
        keep = [i for i, (x, y) in enumerate(zip(sub["start"], sub["end"])) if x >= 0 and y <= duration]
        video = subsample_pad_features(raw, self.max_feats, self.features_dim)
        tok, eos = self.tokenizer, self.tokenizer.eos_token_id
        if keep:
            seq = timed_token_sequence([max(sub["start"][i], 0) for i in keep], [min(sub["end"][i], duration) for i in keep],
                                       [_clean_text(sub["text"][i]) for i in keep], duration, tok, self.num_bins,
                                       self.num_text_tokens, self.max_input_tokens)
            den_in, den_out = span_corrupt(seq, len(tok), self.num_bins, eos, self.noise_density, self.mean_noise_span_length)
        else:
            seq = np.array([eos], dtype=np.int64)
            den_in, den_out = np.array([0], dtype=np.int64), seq
        f = torch.from_numpy
        return {"video_id": video_id, "duration": duration, "video": video, "output_tokens": f(seq),
                "denoising_input_tokens": f(den_in), "denoising_output_tokens": f(den_out)}


_TOKEN_KEYS = ("input_tokens", "output_tokens", "denoising_input_tokens", "denoising_output_tokens")


def collate_dvc(batch: List[dict], pad_to: Optional[Dict[str, int]] = None, pin_memory: bool = False) -> dict:
    """densevideocaptioning_collate_fn (dvc_dataset.py:168-208): zero-pad every token field to the longest row.
    pad_to={"input_tokens": 1000, ...} pads to FIXED lengths instead (never truncates), which keeps the shapes of the
    train step constant so that its CUDA graph can be replayed; pin_memory returns page-locked tensors for async H2D."""
    bs = len(batch)
    out = {"video_id": [b["video_id"] for b in batch], "duration": [b["duration"] for b in batch]}
    video = torch.stack([b["video"] for b in batch])
    out["video"] = video.pin_memory() if pin_memory else video
    for key in _TOKEN_KEYS:
        if key not in batch[0]:
            continue                 # yt_collate_fn (yt_dataset.py:134-165): the pretraining samples carry no "input_tokens"
        rows = [b[key] for b in batch]
        width = max(len(r) for r in rows)
        if pad_to and key in pad_to:
            width = max(width, pad_to[key])
        t = torch.zeros(bs, width, dtype=torch.int64, pin_memory=pin_memory)
        for i, r in enumerate(rows):
            t[i, :len(r)] = r
        out[key] = t
    return out


class PinnedBatcher:
    """Fixed-shape, page-locked staging buffers reused for every batch (no per-batch allocation or pinning): fill() writes
    the samples into them and returns the views `GraphedTrainStep` copies to the device."""

    def __init__(self, batch_size: int, max_feats=100, features_dim=768, max_input_tokens=1000, max_output_tokens=256):
        pin = torch.cuda.is_available()
        self.video = torch.zeros(batch_size, max_feats, features_dim, pin_memory=pin)
        self.tokens = {
            "input_tokens": torch.zeros(batch_size, max_input_tokens, dtype=torch.int64, pin_memory=pin),
            "output_tokens": torch.zeros(batch_size, max_output_tokens, dtype=torch.int64, pin_memory=pin),
            "denoising_input_tokens": torch.zeros(batch_size, max_input_tokens, dtype=torch.int64, pin_memory=pin),
            "denoising_output_tokens": torch.zeros(batch_size, max_input_tokens, dtype=torch.int64, pin_memory=pin),
        }

    def fill(self, batch: List[dict]) -> dict:
        out = {"video_id": [b["video_id"] for b in batch], "duration": [b["duration"] for b in batch], "video": self.video}
        for i, b in enumerate(batch):
            self.video[i].copy_(b["video"])
        for key, buf in self.tokens.items():
            buf.zero_()
            for i, b in enumerate(batch):
                r = b[key]
                assert len(r) <= buf.shape[1], (key, len(r), buf.shape[1])
                buf[i, :len(r)] = r
            out[key] = buf
        return out
