"""Deterministic random initialisation in the reference's scheme.

T5: modeling_t5.py:797-839 (_init_weights, factor=1.0); time-token rows: nn.Embedding default N(0,1)
(HF-4.28 resize_token_embeddings); ViT: model/vit.py:98-111 (trunc_normal pos_embed std .02,
xavier_uniform Linear weights, bias ~ N(0,1e-6), LayerNorm 1/0).  Used where the reference would call
T5ForConditionalGeneration.from_pretrained (no checkpoint files exist offline).
"""
from __future__ import annotations

import math

import torch

from .config import param_shapes


def init_state_dict(cfg: dict, seed: int = 0, emb_std: float = 1.0):
    g = torch.Generator().manual_seed(seed)
    d, dkv, dff, H = cfg["d_model"], cfg["d_kv"], cfg["d_ff"], cfg["num_heads"]
    sd = {}
    for name, shape in param_shapes(cfg):
        t = torch.empty(shape, dtype=torch.float32)
        leaf = name.rsplit(".", 2)
        if name == "t5_model.shared.weight":
            t.normal_(0.0, emb_std, generator=g)
        elif name.endswith("layer_norm.weight") or (name.startswith("visual_encoder") and "norm" in name and name.endswith("weight")):
            t.fill_(1.0)
        elif name.startswith("visual_encoder") and "norm" in name and name.endswith("bias"):
            t.zero_()
        elif name.endswith("relative_attention_bias.weight"):
            t.normal_(0.0, d ** -0.5, generator=g)
        elif ".q.weight" in name:
            t.normal_(0.0, (d * dkv) ** -0.5, generator=g)
        elif ".k.weight" in name or ".v.weight" in name:
            t.normal_(0.0, d ** -0.5, generator=g)
        elif ".o.weight" in name:
            t.normal_(0.0, (H * dkv) ** -0.5, generator=g)
        elif ".wi.weight" in name:
            t.normal_(0.0, d ** -0.5, generator=g)
        elif ".wo.weight" in name:
            t.normal_(0.0, dff ** -0.5, generator=g)
        elif name == "visual_encoder.pos_embed":
            t.normal_(0.0, 0.02, generator=g).clamp_(-0.04, 0.04)
        elif name.endswith(".bias"):
            t.normal_(0.0, 1e-6, generator=g)
        elif name.endswith(".weight") and t.dim() == 2:  # ViT / proj_v2t Linear: xavier_uniform
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t.uniform_(-bound, bound, generator=g)
        else:
            raise KeyError(name)
        del leaf
        sd[name] = t
    return sd
