"""TEST INFRASTRUCTURE — shimmed import of the *unmodified* reference Vid2Seq.

This file loads /root/reference/model/{modeling_t5,vit,vid2seq}.py by path (no
copy) so the reference's own code can be executed in this container, which has
transformers 5.5 / torch 2.11 instead of the pinned 4.28 / 1.13 (SURVEY.md F8,
§8c).  It exists to (a) pin oracle/vid2seq_oracle.py against the real reference
and (b) mint the golden vectors under tests/golden/ (oracle/make_golden.py).

It cannot travel: /root/reference does not exist on the GPU box.  Nothing in
vidchapters_b200/ imports it.  `available()` says whether it can be used.

Shim items (each only patches third-party glue, never reference arithmetic):
  1. transformers.pytorch_utils.find_pruneable_heads_and_indices  (modeling_t5.py:38; used by prune_heads only)
  2. transformers.utils.model_parallel_utils                      (modeling_t5.py:48; used by parallelize only)
  3. PreTrainedModel.get_head_mask -> [None]*n                    (modeling_t5.py:1010-1011)
  4. T5ForConditionalGeneration.from_pretrained -> seeded random init from a T5Config (vid2seq.py:37)
  5. stub tokenizer (len = 32100 + num_bins, pad 0, eos 1)        (vid2seq.py:39-40,86-88)
  6. re-tie lm_head.weight = shared.weight (4.28 ties it; 5.5 leaves it untied)
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("VIDCHAP_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "vid2seq.py"))


class StubTokenizer:
    """Stands in for T5Tokenizer + added <time=i> tokens (vid2seq.py:10-18)."""

    pad_token_id = 0
    eos_token_id = 1

    def __init__(self, base_vocab: int = 32100, num_bins: int = 100):
        self._n = base_vocab + num_bins

    def __len__(self):
        return self._n

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in ids.tolist():
            toks = [t for t in row if not (skip_special_tokens and t in (0, 1))]
            out.append(" ".join(str(t) for t in toks))
        return out


_loaded = None


def load_reference():
    """Returns a namespace with the reference's modules (modeling_t5, vit, vid2seq)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import transformers
    import transformers.pytorch_utils as pu
    from transformers import PreTrainedModel

    # (1)
    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        def find_pruneable_heads_and_indices(*a, **k):  # pragma: no cover - never called on the hot path
            raise NotImplementedError
        pu.find_pruneable_heads_and_indices = find_pruneable_heads_and_indices
    # (2)
    if "transformers.utils.model_parallel_utils" not in sys.modules:
        m = types.ModuleType("transformers.utils.model_parallel_utils")
        m.assert_device_map = lambda *a, **k: None
        m.get_device_map = lambda *a, **k: None
        sys.modules["transformers.utils.model_parallel_utils"] = m
    # (3)
    if not hasattr(PreTrainedModel, "get_head_mask"):
        PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n

    pkg = types.ModuleType("vidchap_ref_model")
    pkg.__path__ = [os.path.join(REF_ROOT, "model")]
    sys.modules["vidchap_ref_model"] = pkg
    mods = {}
    for name in ("modeling_t5", "vit", "vid2seq"):
        full = "vidchap_ref_model." + name
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF_ROOT, "model", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
        setattr(pkg, name, mod)
    _loaded = types.SimpleNamespace(**mods)
    return _loaded


def t5_config(d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12, vocab_size=32128):
    from transformers import T5Config

    return T5Config(
        vocab_size=vocab_size, d_model=d_model, d_kv=d_kv, d_ff=d_ff, num_layers=num_layers,
        num_decoder_layers=num_layers, num_heads=num_heads, feed_forward_proj="relu",
        relative_attention_num_buckets=32, relative_attention_max_distance=128,
        decoder_start_token_id=0, pad_token_id=0, eos_token_id=1, tie_word_embeddings=True,
        dropout_rate=0.1, layer_norm_epsilon=1e-6, initializer_factor=1.0,
    )


def build_reference_vid2seq(cfg: dict, *, vis_drop=0.0, enc_drop=0.0, dec_drop=0.0, label_smoothing=0.1, seed=0,
                            use_video=True, use_speech=True):
    """Builds the reference Vid2Seq (vid2seq.py:20-56) with seeded random init at `cfg` shapes.

    cfg keys: d_model,d_kv,d_ff,num_layers,num_heads,base_vocab,num_bins,num_features,
              embed_dim,depth,heads,mlp_dim.
    """
    ref = load_reference()
    T5 = ref.modeling_t5.T5ForConditionalGeneration
    tok = StubTokenizer(cfg["base_vocab"], cfg["num_bins"])
    conf = t5_config(cfg["d_model"], cfg["d_kv"], cfg["d_ff"], cfg["num_layers"], cfg["num_heads"],
                     vocab_size=cfg["base_vocab"] + 28)

    # (4) from_pretrained -> random init, same ctor kwargs the reference passes (vid2seq.py:37-38)
    def _from_pretrained(cls, encoder_dropout=0.0, decoder_dropout=0.1, label_smoothing=0.1,
                         pretrained_model_name_or_path=None, local_files_only=True, is_gated_act=False, **kw):
        import copy
        torch.manual_seed(seed)
        m = cls(copy.deepcopy(conf), encoder_dropout=encoder_dropout, decoder_dropout=decoder_dropout,
                label_smoothing=label_smoothing, is_gated_act=is_gated_act)
        return m

    orig = T5.__dict__.get("from_pretrained")
    T5.from_pretrained = classmethod(_from_pretrained)
    orig_resize = T5.resize_token_embeddings

    def _resize(self, n):
        # HF-4.28 semantics: keep old rows, new rows ~ N(0,1) (nn.Embedding default), re-tie.
        old = self.shared.weight.data
        new = torch.nn.Embedding(n, old.shape[1])
        k = min(n, old.shape[0])
        new.weight.data[:k] = old[:k]
        self.shared = new
        self.encoder.embed_tokens = new
        self.decoder.embed_tokens = new
        self.lm_head = torch.nn.Linear(old.shape[1], n, bias=False)
        self.lm_head.weight = new.weight  # (6) tie
        self.config.vocab_size = n
        return new

    T5.resize_token_embeddings = _resize
    try:
        torch.manual_seed(seed + 1)
        model = ref.vid2seq.Vid2Seq(
            "t5-base", num_features=cfg["num_features"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
            heads=cfg["heads"], mlp_dim=cfg["mlp_dim"], vis_drop=vis_drop, tokenizer=tok, enc_drop=enc_drop,
            dec_drop=dec_drop, use_speech=use_speech, use_video=use_video, num_bins=cfg["num_bins"],
            label_smoothing=label_smoothing)
    finally:
        T5.resize_token_embeddings = orig_resize
        if orig is not None:
            T5.from_pretrained = orig
        else:
            del T5.from_pretrained
    assert model.t5_model.lm_head.weight is model.t5_model.shared.weight
    return model
