"""Host input pipeline vs the reference's dataset code: bit-exact token streams, features and batches.
Golden: tests/golden/dvc_pipeline.pt, minted by oracle/make_golden_data.py from dataset/dvc_dataset.py + util/t5.py of the
real reference on the synthetic dataset of tests/data_fixture.py; where /root/reference exists the comparison is also
made live (and with another seed)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from data_fixture import HFStubTokenizer, write_dataset  # noqa: E402

from vidchapters_b200 import data as D  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dvc_pipeline.pt")
TOKEN_KEYS = ("input_tokens", "output_tokens", "denoising_input_tokens", "denoising_output_tokens")


def ours(root, seed, batch_size):
    js, feats, subs = write_dataset(root)
    ds = D.DenseVideoCaptioningDataset(js, feats, tokenizer=HFStubTokenizer(), subtitles_path=subs)
    np.random.seed(seed)
    samples = [ds[i] for i in range(len(ds))]
    batches = [D.collate_dvc(samples[i:i + batch_size]) for i in range(0, len(samples), batch_size)]
    return ds, samples, batches


def test_pipeline_matches_reference_golden():
    fx = torch.load(GOLD, weights_only=False)
    with tempfile.TemporaryDirectory() as d:
        ds, samples, batches = ours(d, fx["seed"], fx["batch_size"])
    assert len(samples) == len(fx["samples"]) == 12
    for s, r in zip(samples, fx["samples"]):
        assert s["video_id"] == r["video_id"] and s["duration"] == r["duration"]
        for k in TOKEN_KEYS:
            assert s[k].dtype == torch.int64 and torch.equal(s[k], r[k]), (s["video_id"], k)
        assert s["video"].shape == (100, 768) and s["video"].dtype == torch.float32
        assert s["video"].double().sum().item() == r["video_sum"] and torch.equal(s["video"][:3, :8], r["video_head"])
    for b, r in zip(batches, fx["batches"]):
        assert b["video_id"] == r["video_id"] and tuple(b["video"].shape) == r["video_shape"]
        for k in TOKEN_KEYS:
            assert torch.equal(b[k], r[k]), k
    # the edge cases the fixture plants really occurred
    lens = {k: [len(s[k]) for s in samples] for k in TOKEN_KEYS}
    assert max(lens["input_tokens"]) == 1000 and max(lens["output_tokens"]) == 256          # both truncations
    assert sum(1 for s in samples if len(s["input_tokens"]) == 1) >= 3                        # no / unusable ASR
    assert all(s["input_tokens"][-1] == 1 and s["output_tokens"][-1] == 1 for s in samples)  # eos


@pytest.mark.skipif(not os.path.isdir("/root/reference/dataset"), reason="reference tree not present")
def test_pipeline_matches_live_reference_other_seed():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle"))
    import make_golden_data as G
    with tempfile.TemporaryDirectory() as d:
        ref_s, ref_b = G.reference_outputs(d, seed=7, batch_size=5)
    with tempfile.TemporaryDirectory() as d:
        _, my_s, my_b = ours(d, 7, 5)
    for s, r in zip(my_s, ref_s):
        assert torch.equal(s["video"], r["video"])
        for k in TOKEN_KEYS:
            assert torch.equal(s[k], r[k]), (s["video_id"], k)
    for b, r in zip(my_b, ref_b):
        assert torch.equal(b["video"], r["video"])
        for k in TOKEN_KEYS:
            assert torch.equal(b[k], r[k])


def test_span_corruption_properties():
    """Size-independent properties of T5 span corruption (util/t5.py): noise count and span count are deterministic,
    input and target partition the sequence, sentinels descend from vocab - num_bins - 1."""
    rng = np.random.RandomState(3)
    for n in (2, 3, 17, 64, 999, 1000):
        toks = np.arange(100, 100 + n, dtype=np.int64)
        mask = D.random_spans_noise_mask(n, 0.25, 5, rng)
        k = min(max(int(np.round(n * 0.25)), 1), n - 1)
        assert mask.dtype == bool and mask.shape == (n,) and mask.sum() == k and not mask[0]
        spans = int((mask & ~np.concatenate([[False], mask[:-1]])).sum())
        assert spans == max(int(np.round(k / 5)), 1)
        a, b = D.span_corrupt(toks, 32200, 100, 1, rng=np.random.RandomState(n))
        assert a[-1] == 1 and b[-1] == 1
        body = np.concatenate([a[:-1], b[:-1]])
        assert sorted(body[body < 32000].tolist()) == toks.tolist()          # every token exactly once, in one of the two
        sent_a, sent_b = a[a >= 32000], b[(b >= 32000)]
        assert sent_a.tolist() == list(range(32099, 32099 - len(sent_a), -1)) and sent_b.tolist() == sent_a.tolist()


def test_fixed_shape_and_pinned_batches():
    with tempfile.TemporaryDirectory() as d:
        ds, samples, _ = ours(d, 1, 4)
    pad = {"input_tokens": 1000, "output_tokens": 256, "denoising_input_tokens": 1000, "denoising_output_tokens": 1000}
    b = D.collate_dvc(samples[:4], pad_to=pad)
    assert b["input_tokens"].shape == (4, 1000) and b["output_tokens"].shape == (4, 256)
    free = D.collate_dvc(samples[:4])
    for k in TOKEN_KEYS:
        w = free[k].shape[1]
        assert torch.equal(b[k][:, :w], free[k]) and int(b[k][:, w:].abs().sum()) == 0
    pb = D.PinnedBatcher(4)
    x = pb.fill(samples[4:8])
    y = D.collate_dvc(samples[4:8], pad_to=pad)
    assert torch.equal(x["video"], y["video"]) and all(torch.equal(x[k], y[k]) for k in TOKEN_KEYS)
    z = pb.fill(samples[:4])                                   # buffers are reused: stale rows must be cleared
    assert all(torch.equal(z[k], b[k]) for k in TOKEN_KEYS)


def test_yt_pretraining_pipeline_matches_reference_golden_and_live():
    """dataset/yt_dataset.py (pretraining): golden minted from the reference, plus the live reference where present."""
    from data_fixture import write_yt_dataset
    fx = torch.load(os.path.join(os.path.dirname(GOLD), "yt_pipeline.pt"), weights_only=False)
    keys = ("output_tokens", "denoising_input_tokens", "denoising_output_tokens")

    def mine(seed, bs):
        with tempfile.TemporaryDirectory() as d:
            csv, feats, subs = write_yt_dataset(d)
            ds = D.YTDataset(csv, feats, subs, tokenizer=HFStubTokenizer())
            np.random.seed(seed)
            samples = [ds[i] for i in range(len(ds))]
        return samples, [D.collate_dvc(samples[i:i + bs]) for i in range(0, len(samples), bs)]

    samples, batches = mine(fx["seed"], fx["batch_size"])
    assert len(samples) == len(fx["samples"]) == 10
    for s, r in zip(samples, fx["samples"]):
        assert s["video_id"] == r["video_id"] and s["duration"] == r["duration"] and "input_tokens" not in s
        assert s["video"].double().sum().item() == r["video_sum"]
        for k in keys:
            assert torch.equal(s[k], r[k]), (s["video_id"], k)
    for b, r in zip(batches, fx["batches"]):
        assert tuple(b["video"].shape) == r["video_shape"] and set(b) == set(r) - {"video_shape"} | {"video"}
        for k in keys:
            assert torch.equal(b[k], r[k]), k
    assert max(len(s["output_tokens"]) for s in samples) == 1000 and min(len(s["output_tokens"]) for s in samples) == 1
    if os.path.isdir("/root/reference/dataset"):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle"))
        import make_golden_data as G
        with tempfile.TemporaryDirectory() as d:
            ref_s, ref_b = G.reference_yt_outputs(d, seed=11, batch_size=3)
        my_s, my_b = mine(11, 3)
        for s, r in zip(my_s, ref_s):
            assert torch.equal(s["video"], r["video"]) and all(torch.equal(s[k], r[k]) for k in keys)
        for b, r in zip(my_b, ref_b):
            assert torch.equal(b["video"], r["video"]) and all(torch.equal(b[k], r[k]) for k in keys)
