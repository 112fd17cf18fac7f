"""TEST INFRASTRUCTURE — mints tests/golden/*.pt from the REAL reference (through oracle/ref_shim.py).

Run here (the container that has /root/reference):  python -m oracle.make_golden
Each fixture holds the inputs and the reference's outputs for one (config, batch); the weights are NOT stored —
they are regenerated bit-identically by vidchapters_b200.init.init_state_dict(cfg, seed) (CPU torch.Generator) and
loaded into the reference with load_state_dict, dropout 0 (SURVEY §8c parity protocol).
Stored: loss, logits (full for tiny configs; slices + argmax for t5-base), video-encoder output, per-parameter gradient
norms, a few gradient slices, and the parameters after one dvc.py:112-126 step for selected tensors.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from vidchapters_b200.config import T5_BASE, TINY, TINY_PROJ  # noqa: E402
from vidchapters_b200.init import init_state_dict  # noqa: E402

CASES = [
    ("tiny", dict(TINY, num_features=10), dict(B=2, T=10, L=24, S=12, seed=1), True),
    ("tiny_proj", dict(TINY_PROJ), dict(B=2, T=7, L=40, S=20, seed=2), True),
    ("tiny_long", dict(TINY), dict(B=2, T=100, L=300, S=140, seed=3), True),
    # BASELINE.json configs[0]: t5-base, 1 video, 10 frames, 64 ASR tokens, 32 target tokens
    ("t5base_cfg1", dict(T5_BASE), dict(B=1, T=10, L=64, S=32, seed=1), False),
    # vc.py path (vc.py:26-86,280,299-312): no time tokens (num_bins=0), a vocabulary that is NOT a multiple of 8 (like
    # 32100), a ragged `padding="longest"` batch whose lengths are multiples of nothing, clips shorter than max_feats
    ("tiny_vc", dict(TINY, num_features=10, num_bins=0, base_vocab=1012), dict(B=3, T=7, L=37, S=19, seed=4), True),
]


def make_batch(cfg, B, T, L, S, seed):
    g = torch.Generator().manual_seed(seed)
    V = cfg["base_vocab"] + cfg["num_bins"]
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, V, (B, L), generator=g)
    out = torch.randint(2, cfg["base_vocab"], (B, S), generator=g)
    # time tokens in the targets (argmax over the time-token range is a parity check), eos, ragged padding
    if cfg["num_bins"]:
        out[:, 0::5] = torch.randint(cfg["base_vocab"], V, out[:, 0::5].shape, generator=g)
    for b in range(B):
        li = int(torch.randint(L // 2, L + 1, (1,), generator=g)) if b else L
        lo = int(torch.randint(S // 2, S + 1, (1,), generator=g)) if b else S
        inp[b, li - 1] = 1
        inp[b, li:] = 0
        out[b, lo - 1] = 1
        out[b, lo:] = 0
    return video, inp, out


def main(only=None):
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, cfg, bs, full in CASES:
        if only and name not in only:
            continue
        torch.manual_seed(0)
        model = ref_shim.build_reference_vid2seq(cfg, vis_drop=0.0, enc_drop=0.0, dec_drop=0.0)
        sd = init_state_dict(cfg, 0)
        fullsd = dict(sd)
        for k in ("t5_model.encoder.embed_tokens.weight", "t5_model.decoder.embed_tokens.weight", "t5_model.lm_head.weight"):
            fullsd[k] = sd["t5_model.shared.weight"]
        model.load_state_dict(fullsd, strict=True)
        model.train()
        video, inp, out = make_batch(cfg, **bs)
        it = {"input_ids": inp, "attention_mask": inp != 0}
        ot = {"input_ids": out, "attention_mask": out != 0}
        loss_dict, vd = model(video, it, ot)
        loss = loss_dict["loss"]
        model.zero_grad()
        loss.backward()
        # logits the way SURVEY F4 prescribes: call t5_model with the same kwargs as vid2seq.py:89-95
        with torch.no_grad():
            from transformers.modeling_outputs import BaseModelOutput
            text = model.t5_model.encoder.embed_tokens(inp)
            enc = model.t5_model.encoder(attention_mask=inp != 0, inputs_embeds=text)
            mem = torch.cat([vd["video"], enc.last_hidden_state], dim=1)
            atts = torch.cat([vd["atts_vis"], (inp != 0).long()], dim=1)
            targets = out.masked_fill(out == 0, -100)
            o = model.t5_model(encoder_outputs=BaseModelOutput(last_hidden_state=mem), attention_mask=atts,
                               decoder_attention_mask=out != 0, return_dict=True, labels=targets)
            logits = o.logits
        named = dict(model.named_parameters())
        fx = dict(name=name, cfg=cfg, batch=bs, video=video, input_ids=inp, output_ids=out, loss=loss.detach().clone(),
                  video_out=vd["video"].detach().clone(), memory=mem.clone(),
                  grad_norms={n: p.grad.norm().item() for n, p in named.items()})
        V0 = cfg["base_vocab"]
        fx["logits_argmax"] = logits.argmax(-1)
        if cfg["num_bins"]:
            fx["time_argmax"] = logits[..., V0:].argmax(-1)
        if full:
            fx["logits"] = logits.clone()
            fx["grads"] = {n: p.grad.clone() for n, p in named.items()
                           if p.numel() <= 1 << 16 or n.endswith("block.0.layer.0.SelfAttention.q.weight")}
        else:
            fx["logits_time"] = logits[..., V0:].clone()
            fx["logits_head"] = logits[..., :512].clone()
            fx["logits_norm"] = logits.norm().item()
            fx["grads"] = {n: p.grad.clone() for n, p in named.items() if p.numel() <= 1 << 12}
        # one optimiser step exactly as dvc.py:112-126 (README runs use clip 0.1)
        opt = torch.optim.Adam(model.parameters(), lr=3e-4, betas=(0.9, 0.999), weight_decay=0)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        with torch.no_grad():
            nb = cfg["num_bins"]
            for w in ((model.t5_model.shared.weight, model.t5_model.lm_head.weight) if nb else ()):  # vc.py: no renorm
                frozen = torch.norm(w[:-nb, :], dim=1).mean(0)
                w[-nb:, :].div_(torch.norm(w[-nb:, :], dim=1).mean(0) / frozen)
        fx["after_step"] = {
            "time_rows": model.t5_model.shared.weight[-max(cfg["num_bins"], 1):].detach().clone(),
            "enc_ln": named["t5_model.encoder.final_layer_norm.weight"].detach().clone(),
            "rel_bias": named["t5_model.encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"].detach().clone(),
            "vit_norm_b": named["visual_encoder.norm.bias"].detach().clone(),
        }
        path = os.path.join(ROOT, "tests", "golden", name + ".pt")
        torch.save(fx, path)
        print(name, "loss", float(loss), "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main(set(sys.argv[1:]))   # python -m oracle.make_golden [case names]; no names = all
