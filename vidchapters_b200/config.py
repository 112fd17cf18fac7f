"""Model shape configs for the Vid2Seq hot path (reference: model/vid2seq.py:21-56, args.py:217-305).

Keys: T5 side d_model,d_kv,d_ff,num_layers,num_heads,base_vocab(+num_bins time tokens);
      visual encoder side num_features,embed_dim,depth,heads,mlp_dim (model/vit.py:83-96).
"""
from __future__ import annotations

T5_BASE = dict(name="t5-base", d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12, base_vocab=32100,
               num_bins=100, num_features=100, embed_dim=768, depth=12, heads=12, mlp_dim=2048)
T5_LARGE = dict(name="t5-large", d_model=1024, d_kv=64, d_ff=4096, num_layers=24, num_heads=16, base_vocab=32100,
                num_bins=100, num_features=100, embed_dim=768, depth=12, heads=12, mlp_dim=2048)
# Reduced-depth configs for fast parity tests (same per-layer shapes as t5-base; small vocab).
TINY = dict(name="tiny", d_model=768, d_kv=64, d_ff=3072, num_layers=2, num_heads=12, base_vocab=1000,
            num_bins=100, num_features=100, embed_dim=768, depth=2, heads=12, mlp_dim=2048)
# d_model != 768 exercises proj_v2t (vid2seq.py:54-56), ragged head count.
TINY_PROJ = dict(name="tiny-proj", d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, base_vocab=1000,
                 num_bins=100, num_features=100, embed_dim=768, depth=1, heads=12, mlp_dim=2048)

CONFIGS = {c["name"]: c for c in (T5_BASE, T5_LARGE, TINY, TINY_PROJ)}


def vocab_size(cfg: dict) -> int:
    return cfg["base_vocab"] + cfg["num_bins"]


def param_shapes(cfg: dict):
    """Ordered (name, shape) list in the reference's state-dict key space.

    Order is chosen so that q,k,v (and cross-attn k,v) weights are adjacent in the flat parameter
    buffer: a [3d,d] (resp. [2d,d]) fused-QKV view then exists with no packing step.
    """
    d, dkv, dff, H = cfg["d_model"], cfg["d_kv"], cfg["d_ff"], cfg["num_heads"]
    inner = H * dkv
    V = vocab_size(cfg)
    C, mlp = cfg["embed_dim"], cfg["mlp_dim"]
    out = [("t5_model.shared.weight", (V, d))]
    for stack in ("encoder", "decoder"):
        for i in range(cfg["num_layers"]):
            p = f"t5_model.{stack}.block.{i}.layer."
            out += [(p + "0.SelfAttention.q.weight", (inner, d)), (p + "0.SelfAttention.k.weight", (inner, d)),
                    (p + "0.SelfAttention.v.weight", (inner, d)), (p + "0.SelfAttention.o.weight", (d, inner))]
            if i == 0:
                out.append((p + "0.SelfAttention.relative_attention_bias.weight", (32, H)))
            out.append((p + "0.layer_norm.weight", (d,)))
            ff = 1
            if stack == "decoder":
                out += [(p + "1.EncDecAttention.q.weight", (inner, d)), (p + "1.EncDecAttention.k.weight", (inner, d)),
                        (p + "1.EncDecAttention.v.weight", (inner, d)), (p + "1.EncDecAttention.o.weight", (d, inner)),
                        (p + "1.layer_norm.weight", (d,))]
                ff = 2
            out += [(p + f"{ff}.DenseReluDense.wi.weight", (dff, d)), (p + f"{ff}.DenseReluDense.wo.weight", (d, dff)),
                    (p + f"{ff}.layer_norm.weight", (d,))]
        out.append((f"t5_model.{stack}.final_layer_norm.weight", (d,)))
    out.append(("visual_encoder.pos_embed", (1, cfg["num_features"], C)))
    for i in range(cfg["depth"]):
        p = f"visual_encoder.blocks.{i}."
        out += [(p + "norm1.weight", (C,)), (p + "norm1.bias", (C,)),
                (p + "attn.qkv.weight", (3 * C, C)), (p + "attn.qkv.bias", (3 * C,)),
                (p + "attn.proj.weight", (C, C)), (p + "attn.proj.bias", (C,)),
                (p + "norm2.weight", (C,)), (p + "norm2.bias", (C,)),
                (p + "mlp.fc1.weight", (mlp, C)), (p + "mlp.fc1.bias", (mlp,)),
                (p + "mlp.fc2.weight", (C, mlp)), (p + "mlp.fc2.bias", (C,))]
    out += [("visual_encoder.norm.weight", (C,)), ("visual_encoder.norm.bias", (C,))]
    if d != 768:
        out += [("proj_v2t.weight", (d, 768)), ("proj_v2t.bias", (d,))]
    return out


def layout_order(cfg: dict):
    """param_shapes(cfg) in the order the engine lays parameters out in its flat buffers: identical, except that the
    decoder's cross-attention k,v weights of ALL layers form one contiguous group (right before the decoder's final norm),
    so that a single [num_layers * 2 * inner, d] matrix view exists — the K/V projections of every layer read the same
    encoder memory and can run as ONE GEMM (engine.fuse_cross_kv).  Initialisation keeps following param_shapes' order."""
    shapes = param_shapes(cfg)
    is_ckv = lambda n: ".layer.1.EncDecAttention.k.weight" in n or ".layer.1.EncDecAttention.v.weight" in n
    group = [e for e in shapes if is_ckv(e[0])]
    out = []
    for e in shapes:
        if is_ckv(e[0]):
            continue
        if e[0] == "t5_model.decoder.final_layer_norm.weight":
            out += group
        out.append(e)
    return out
