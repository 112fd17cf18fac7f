"""TEST INFRASTRUCTURE — pins oracle.beam_search_decode / greedy_decode against a real HuggingFace `generate`.

The reference delegates decoding to transformers==4.28.0 `GenerationMixin.generate` (model/vid2seq.py:150-162), which is
not vendored and not installable here.  The container has transformers 5.5, whose stock `T5ForConditionalGeneration`
computes the same decoder arithmetic as the reference's fork at dropout 0 (tied LM head scaled by d_model**-0.5) and
whose `generate(num_beams=k, do_sample=False, early_stopping=False)` implements the same published beam search.  This
script trains a very small T5 decoder (so that hypotheses really end with eos at different lengths and ranks), runs the
stock HF `generate` on it and stores the weights, inputs and HF's token ids in tests/golden/beam_hf.pt;
tests/test_oracle_cpu.py checks the oracle against the stored ids everywhere and against a live HF run where
transformers imports.   python -m oracle.make_golden_beam
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import vid2seq_oracle as O  # noqa: E402
from vidchapters_b200.init import init_state_dict  # noqa: E402

CFG = dict(name="beam-mini", d_model=64, d_kv=64, d_ff=128, num_layers=2, num_heads=2, base_vocab=112, num_bins=8,
           num_features=6, embed_dim=768, depth=1, heads=12, mlp_dim=64)
CASES = [(1, 1.0, 14), (2, 1.0, 14), (4, 1.0, 14), (4, 0.6, 14), (4, 2.0, 14), (3, 1.0, 6), (8, 1.0, 14)]


def hf_model(cfg, sd):
    from transformers import T5Config, T5ForConditionalGeneration
    V = cfg["base_vocab"] + cfg["num_bins"]
    conf = T5Config(vocab_size=V, d_model=cfg["d_model"], d_kv=cfg["d_kv"], d_ff=cfg["d_ff"], num_layers=cfg["num_layers"],
                    num_decoder_layers=cfg["num_layers"], num_heads=cfg["num_heads"], feed_forward_proj="relu",
                    relative_attention_num_buckets=32, relative_attention_max_distance=128, decoder_start_token_id=0,
                    pad_token_id=0, eos_token_id=1, tie_word_embeddings=True, dropout_rate=0.0, layer_norm_epsilon=1e-6)
    hf = T5ForConditionalGeneration(conf).eval()
    hsd = {k[len("t5_model."):]: v.detach().clone() for k, v in sd.items() if k.startswith("t5_model.")}
    for alias in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight"):
        hsd[alias] = hsd["shared.weight"]
    missing, unexpected = hf.load_state_dict(hsd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return hf


def hf_generate(hf, memory, mask, nb, lp, max_new):
    from transformers.modeling_outputs import BaseModelOutput
    with torch.no_grad():
        return hf.generate(encoder_outputs=BaseModelOutput(last_hidden_state=memory), attention_mask=mask, num_beams=nb,
                           max_new_tokens=max_new, min_length=1, do_sample=False, length_penalty=lp, early_stopping=False,
                           num_return_sequences=1, repetition_penalty=1.0)


def main():
    cfg = CFG
    torch.manual_seed(0)
    sd = {k: v.requires_grad_(True) for k, v in init_state_dict(cfg, 1, emb_std=0.5).items() if k.startswith("t5_model.")}
    g = torch.Generator().manual_seed(2)
    B, E, S = 6, 9, 12
    V = cfg["base_vocab"] + cfg["num_bins"]
    memory = torch.randn(B, E, cfg["d_model"], generator=g)
    mask = torch.ones(B, E, dtype=torch.long)
    mask[2, -3:] = 0
    # targets of different lengths (eos at positions 2..10); several samples share prefixes so that beams compete
    tgt = torch.randint(2, V, (B, S), generator=g)
    tgt[1, :3] = tgt[0, :3]
    tgt[4, :5] = tgt[3, :5]
    for b, n in enumerate((3, 6, 9, 5, 8, 11)):
        tgt[b, n - 1] = 1
        tgt[b, n:] = 0
    opt = torch.optim.Adam(list(sd.values()), lr=3e-3)
    ar = O.Arith(False)
    for step in range(int(os.environ.get("BEAM_TRAIN_STEPS", "35"))):          # half-trained on purpose: near-ties between hypotheses, eos at several ranks
        labels = tgt.masked_fill(tgt == 0, -100)
        dec_in = O.shift_right(labels)
        seq = O.t5_decoder(sd, cfg, dec_in, tgt != 0, memory, mask, ar) * (cfg["d_model"] ** -0.5)
        logits = ar.linear(seq, sd["t5_model.shared.weight"])
        loss = torch.nn.functional.cross_entropy(logits.view(-1, V), labels.view(-1), ignore_index=-100)
        opt.zero_grad(); loss.backward(); opt.step()
    print("trained: loss", float(loss))
    sd = {k: v.detach() for k, v in sd.items()}
    hf = hf_model(cfg, sd)
    # perturbed memories: the model is uncertain on them
    mem2 = memory + 0.7 * torch.randn(memory.shape, generator=g)
    cases = []
    for mem_name, mem in (("train", memory), ("perturbed", mem2)):
        for nb, lp, max_new in CASES:
            ids = hf_generate(hf, mem, mask, nb, lp, max_new)
            cases.append(dict(memory=mem_name, num_beams=nb, length_penalty=lp, max_new_tokens=max_new, ids=ids))
            print(mem_name, nb, lp, max_new, ids[:3].tolist())
    import transformers
    path = os.path.join(ROOT, "tests", "golden", "beam_hf.pt")
    torch.save(dict(cfg=cfg, sd=sd, memory={"train": memory, "perturbed": mem2}, mask=mask, cases=cases,
                    transformers_version=transformers.__version__), path)
    print("->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
