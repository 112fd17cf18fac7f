"""Drop-in `Vid2Seq` module: the reference's Python surface over the B200 engine.

Mirrors model/vid2seq.py:20-167 and model/__init__.py:4-19 of the reference: same constructor arguments, same
`forward(video, input_tokenized, output_tokenized) -> ({"loss": loss}, video_dict)`, same attribute names touched from
dvc.py / vc.py (`t5_model.shared.weight`, `t5_model.lm_head.weight`, `t5_tokenizer`, `visual_encoder`, `proj_v2t`,
`use_video`, `use_speech`) and the same state-dict key space (incl. the four aliased embedding keys),
so `dvc.py`'s loop — forward, `loss.backward()`, `clip_grad_norm_`, `optimizer.step()`, time-token renorm — runs
unchanged.  Underneath, forward and backward are single autograd nodes that call the hand-written sm_100a kernels
(vidchapters_b200.engine); there is no CPU / PyTorch fallback: calling the module without a B200 raises.

All parameters are views into ONE flat fp32 buffer (what the fused optimiser and the gradient all-reduce want); their
`.grad`s are views into one flat fp32 gradient buffer.
"""
from __future__ import annotations

import contextlib
import os
import warnings
from typing import Optional

import torch
import torch.nn as nn

from .config import CONFIGS, param_shapes
from .engine import Vid2SeqEngine
from .init import init_state_dict


def _get_tokenizer(tokenizer_path, num_bins=0):
    """Same contract as the reference's model/vid2seq.py:10-18 (T5Tokenizer + <time=i> tokens)."""
    if "t5" in tokenizer_path:
        from transformers import T5Tokenizer
        tokenizer = T5Tokenizer.from_pretrained(tokenizer_path, local_files_only=True)
        if num_bins:
            tokenizer.add_tokens(["<time=" + str(i) + ">" for i in range(num_bins)])
    else:
        raise NotImplementedError(tokenizer_path)
    return tokenizer


def _dev_guard(device):
    """Every engine entry point runs with the MODEL's device current: the C ABI launches on the current device / its
    current stream, and a model on cuda:1 driven while cuda:0 is current would otherwise launch on the wrong GPU."""
    device = torch.device(device)
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


_HF_WEIGHT_FILES = ("model.safetensors", "pytorch_model.bin")


def find_pretrained_t5(t5_path) -> Optional[str]:
    """Weight file of an HF T5 checkpoint directory (what `T5ForConditionalGeneration.from_pretrained(t5_path,
    local_files_only=True)` would read, model/vid2seq.py:37-38), or None."""
    if t5_path is None or not os.path.isdir(str(t5_path)):
        return None
    for f in _HF_WEIGHT_FILES:
        if os.path.isfile(os.path.join(str(t5_path), f)):
            return os.path.join(str(t5_path), f)
    return None


def load_hf_t5_state_dict(path: str) -> dict:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def t5_state_to_vid2seq(hf_sd: dict, cfg: dict, generator: Optional[torch.Generator] = None) -> dict:
    """HF `T5ForConditionalGeneration` state dict -> this module's key space, with the reference's vocabulary surgery
    (model/vid2seq.py:37-40): `from_pretrained` (32128 rows for the released T5s) -> `resize_token_embeddings(len(tok) -
    num_bins)` crops to the tokenizer's 32100 rows -> `resize_token_embeddings(len(tok))` appends num_bins rows, which
    HF-4.28 leaves at nn.Embedding's default N(0,1) init (T5's `_init_weights` has no Embedding branch) ->
    lm_head re-tied to `shared`.  Keys: `shared.weight`, `{encoder,decoder}.block.*`, `*.final_layer_norm.weight` get the
    `t5_model.` prefix; `lm_head.weight` / `*.embed_tokens.weight` are aliases of `shared` (tied) and are dropped, as is
    the cross-attention relative bias some old checkpoints carry."""
    d, V0, nb = cfg["d_model"], cfg["base_vocab"], cfg["num_bins"]
    shared = hf_sd["shared.weight"].float()
    if shared.shape[1] != d:
        raise ValueError(f"checkpoint d_model {shared.shape[1]} != configured {d}")
    if shared.shape[0] < V0:
        raise ValueError(f"checkpoint vocabulary {shared.shape[0]} smaller than the tokenizer's {V0}")
    new_rows = torch.randn(nb, d, generator=generator)            # nn.Embedding default init of the appended rows
    out = {"t5_model.shared.weight": torch.cat([shared[:V0], new_rows], 0)}
    for k, v in hf_sd.items():
        if k.startswith(("encoder.block.", "decoder.block.")) or k.endswith("final_layer_norm.weight"):
            if "EncDecAttention.relative_attention_bias" in k:
                continue
            out["t5_model." + k] = v.float()
    return out


class _Node(nn.Module):
    """Bare container used to reproduce the reference's module tree (and hence its state-dict keys)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-modules of vidchapters_b200.Vid2Seq are parameter containers; call the model itself")


class _Embedding(_Node):
    def forward(self, ids):  # API compatibility only (vid2seq.py:71 calls encoder.embed_tokens); not on the hot path
        return torch.nn.functional.embedding(ids, self.weight)


class _Vid2SeqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, anchor, video, input_ids, input_mask, output_ids, output_mask, cached):
        eng = module.engine
        eng.drop_rates = dict(vis=module.vis_drop, enc=module.enc_drop, dec=module.dec_drop)
        loss, ectx = eng.forward(video, input_ids, input_mask, output_ids, output_mask, video_cached=cached,
                                 training=module.training)
        ctx.module, ctx.ectx = module, ectx
        ctx.set_materialize_grads(False)
        B, T = ectx["B"], ectx["T"]
        vid = ectx["vid_f32"]
        vid = vid.view(B, T, -1) if vid is not None else loss.new_zeros(0)
        return loss.view(()), vid

    @staticmethod
    def backward(ctx, grad_loss, grad_vid):
        module = ctx.module
        with _dev_guard(module._flat.device):
            module._begin_backward()
            if grad_loss is None:
                grad_loss = torch.zeros(1, device=module._flat.device)
            dvideo = module.engine.backward(ctx.ectx, grad_loss, grad_vid)
            ctx.ectx = None
            module._end_backward()
        return None, None, dvideo, None, None, None, None, None


class Vid2Seq(nn.Module):
    def __init__(self, t5_path, num_features=100, embed_dim=768, depth=12, heads=12, mlp_dim=2048, vis_drop=0.0,
                 tokenizer=None, enc_drop=0.0, dec_drop=0.1, use_speech=True, use_video=True, num_bins=100,
                 label_smoothing=0.1, *, t5_config: Optional[dict] = None, seed: int = 0, ops=None,
                 pretrained: Optional[bool] = None):
        """`t5_path`: as in the reference (model/vid2seq.py:37-38) the directory of an HF T5 checkpoint
        (`$TRANSFORMERS_CACHE/t5-base`, dvc.py:480).  If it holds `model.safetensors` / `pytorch_model.bin` the T5 weights
        are loaded from it with the reference's vocabulary surgery (32128 -> 32100 -> +num_bins rows ~ N(0,1)); the
        visual encoder is seeded-random in the reference's init scheme, like the reference's.  "t5-base" / "t5-large"
        in the path selects the shape.  `pretrained`: None = load if the files exist, otherwise WARN that the T5 is
        randomly initialised (the reference would raise: no silent random T5); True = raise if they are missing; False =
        seeded random init without a warning (benchmarks, tests).  `t5_config` (explicit shape dict; implies
        pretrained=False unless files exist), `seed`, `ops` (injected op table: tests use the torch oracle table on CPU)
        are extensions."""
        super().__init__()
        if "v1_1" in str(t5_path):
            raise NotImplementedError("gated-activation T5 v1.1 checkpoints (is_gated_act, model/vid2seq.py:38) are not "
                                      "supported by the B200 path: t5-base / t5-large use ReLU feed-forward layers")
        weight_file = find_pretrained_t5(t5_path)
        if weight_file is None:
            if pretrained:
                raise OSError(f"no T5 checkpoint (model.safetensors / pytorch_model.bin) under {t5_path!r}")
            if pretrained is None and t5_config is None:
                warnings.warn(f"vidchapters_b200.Vid2Seq: no T5 checkpoint found under {t5_path!r}; the T5 weights are "
                              "RANDOMLY INITIALISED (the reference loads pretrained weights here, model/vid2seq.py:37). "
                              "Load a checkpoint with load_state_dict / --load, or pass pretrained=False to silence.",
                              RuntimeWarning, stacklevel=2)
        if t5_config is None:
            key = "t5-large" if "large" in str(t5_path) else "t5-base"
            t5_config = dict(CONFIGS[key])
        cfg = dict(t5_config)
        if tokenizer is not None:
            cfg["base_vocab"] = len(tokenizer) - num_bins
        cfg.update(num_bins=num_bins, num_features=num_features, embed_dim=embed_dim, depth=depth, heads=heads,
                   mlp_dim=mlp_dim)
        self.cfg = cfg
        self.t5_tokenizer = tokenizer
        self.use_speech, self.use_video = use_speech, use_video
        self.vis_drop, self.enc_drop, self.dec_drop = vis_drop, enc_drop, dec_drop
        self.label_smoothing = label_smoothing
        self._ops = ops
        self._engine: Optional[Vid2SeqEngine] = None
        self._shadow_valid = False

        # flat fp32 parameter buffer + the reference's module tree of Parameter views
        probe = Vid2SeqEngine.layout_of(cfg)
        self._layout, total = probe
        self._flat = torch.zeros(total, dtype=torch.float32)
        sd = init_state_dict(cfg, seed)
        self.pretrained_from = None
        if weight_file is not None and pretrained is not False:
            g = torch.Generator().manual_seed(seed)
            loaded = t5_state_to_vid2seq(load_hf_t5_state_dict(weight_file), cfg, g)
            missing = [n for n, _ in param_shapes(cfg) if n.startswith("t5_model.") and n not in loaded]
            if missing:
                raise KeyError(f"T5 checkpoint {weight_file} lacks {missing[:4]} ... ({len(missing)} tensors)")
            for n, t in loaded.items():
                if n in sd:
                    if tuple(t.shape) != tuple(sd[n].shape):
                        raise ValueError(f"{n}: checkpoint shape {tuple(t.shape)} != model shape {tuple(sd[n].shape)}")
                    sd[n] = t
            self.pretrained_from = weight_file
        self._params = {}
        for name, shape in param_shapes(cfg):
            o, shp, n = self._layout[name]
            self._flat[o:o + n].copy_(sd[name].reshape(-1))
            par = nn.Parameter(self._flat[o:o + n].view(shp))
            self._params[name] = par
            self._attach(name, par)
        del sd
        t5 = self.t5_model
        t5.encoder.add_module("embed_tokens", t5.shared)   # aliases, as in the reference (modeling_t5.py:1507-1532)
        t5.decoder.add_module("embed_tokens", t5.shared)
        lm = _Node()
        lm.weight = t5.shared.weight                        # tied lm_head
        t5.add_module("lm_head", lm)
        t5.model_dim = cfg["d_model"]
        if cfg["d_model"] == 768:
            self.proj_v2t = None

    # ------------------------------------------------------------------ module tree
    def _attach(self, name, par):
        parts = name.split(".")
        node = self
        for i, part in enumerate(parts[:-1]):
            nxt = node._modules.get(part)
            if nxt is None:
                nxt = _Embedding() if (part == "shared") else _Node()
                node.add_module(part, nxt)
            node = nxt
        node.register_parameter(parts[-1], par)

    def _apply(self, fn, recurse=True):
        """.to()/.cuda(): move the flat buffer and re-point every Parameter view (keeps them aliased)."""
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise RuntimeError("vidchapters_b200.Vid2Seq keeps fp32 master weights (bf16 compute copies are internal)")
        if new_flat is not self._flat:
            self._flat = new_flat
            for name, par in self._params.items():
                o, shp, n = self._layout[name]
                par.data = self._flat[o:o + n].view(shp)
                par.grad = None
            self._engine = None
            self._shadow_valid = False
        return self

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=False)
        self._shadow_valid = False
        return r

    # ------------------------------------------------------------------ engine
    @property
    def engine(self) -> Vid2SeqEngine:
        if self._engine is None:
            dev = self._flat.device
            ops = self._ops
            if ops is None:
                if dev.type != "cuda":
                    raise RuntimeError("vidchapters_b200.Vid2Seq runs on a B200 only: move it with .to('cuda') first "
                                       "(there is no CPU fallback path)")
                from .ops import CudaOps
                with torch.cuda.device(dev):
                    ops = CudaOps()
            self._engine = Vid2SeqEngine(self.cfg, ops, dev, label_smoothing=self.label_smoothing,
                                         use_video=self.use_video, use_speech=self.use_speech, flat_p=self._flat)
            self._shadow_valid = False
        return self._engine

    def _refresh_shadow(self):
        eng = self.engine
        if not self._shadow_valid:
            eng.sync_bf16()   # parameters may have been changed by load_state_dict / a stock optimiser / dvc.py's renorm
        self._shadow_valid = False  # only the fused optimiser (which maintains the shadow itself) re-validates it

    def _begin_backward(self):
        if self._params["t5_model.shared.weight"].grad is None:
            self.engine.zero_grad()

    def _end_backward(self):
        eng = self.engine
        for name, par in self._params.items():
            if par.grad is None:
                par.grad = eng.g(name)

    # ------------------------------------------------------------------ reference surface
    def forward(self, video, input_tokenized, output_tokenized):
        with _dev_guard(self._flat.device):
            return self._forward(video, input_tokenized, output_tokenized)

    def _forward(self, video, input_tokenized, output_tokenized):
        self._refresh_shadow()
        cached = isinstance(video, dict)
        if self.use_video:
            vid_in = video["video"] if cached else video
        else:
            vid_in = None
        input_ids = input_tokenized["input_ids"] if self.use_speech else None
        input_mask = input_tokenized["attention_mask"] if self.use_speech else None
        out_ids = output_tokenized["input_ids"]
        out_mask = output_tokenized["attention_mask"]
        anchor = self._params["t5_model.shared.weight"]
        loss, vid = _Vid2SeqFn.apply(self, anchor, vid_in, input_ids, input_mask, out_ids, out_mask, cached)
        video_dict = None
        if self.use_video:
            atts = video["atts_vis"] if cached else torch.ones(vid.shape[:2], dtype=torch.long, device=vid.device)
            video_dict = {"video": vid, "atts_vis": atts}
        return {"loss": loss}, video_dict

    @torch.no_grad()
    def forward_logits(self, video, input_tokenized, output_tokenized):
        """Debug path: materialises (B,S,V) logits like `model.t5_model(...).logits` in the reference."""
        with _dev_guard(self._flat.device):
            self._refresh_shadow()
            eng = self.engine
            loss, ectx = eng.forward(video, input_tokenized["input_ids"], input_tokenized["attention_mask"],
                                     output_tokenized["input_ids"], output_tokenized["attention_mask"], want_logits=True)
            B, S = ectx["B"], ectx["S"]
            return loss.view(()), ectx["logits"].reshape(B, S, -1)

    @torch.no_grad()
    def generate(self, video, input_tokenized, use_nucleus_sampling=False, num_beams=4, max_length=256, min_length=1,
                 top_p=0.9, repetition_penalty=1.0, length_penalty=1.0, num_captions=1, temperature=1):
        """Same kwargs as the reference (model/vid2seq.py:100-167), which forwards them to HF-4.28 `generate`:
        greedy (num_beams=1), beam search (2..8 beams; the reference default is 4) and nucleus sampling
        (`use_nucleus_sampling`, i.e. dvc.py's `--num_beams 0`) run on the B200 path with a KV cache and a CUDA-graphed
        decode step; `repetition_penalty`, `min_length`, `temperature`, `top_p`, `length_penalty` and `num_captions`
        (num_return_sequences) follow HF's semantics (processors pinned against stock HF generate in
        tests/test_oracle_cpu.py).  Sampling draws from torch's random stream of the model's device.  Not built (raise):
        beam-sampling (sampling with num_beams > 1) and more than 8 beams."""
        if num_beams > 8 or num_beams < 0 or num_captions < 1:
            raise NotImplementedError("vidchapters_b200.Vid2Seq.generate: 0 <= num_beams <= 8, num_captions >= 1")
        if use_nucleus_sampling and num_beams > 1:
            raise NotImplementedError("beam-sampling (use_nucleus_sampling with num_beams > 1) is not built")
        if not use_nucleus_sampling and num_beams == 0:
            raise ValueError("num_beams=0 selects nucleus sampling in dvc.py (use_nucleus_sampling=True)")
        if not use_nucleus_sampling and num_captions > 1 and (num_beams == 1 or num_captions > num_beams):
            raise ValueError("num_captions > 1 needs beam search (num_captions <= num_beams) or sampling, as in HF generate")
        with _dev_guard(self._flat.device):
            self._refresh_shadow()
            eng = self.engine
            ids = input_tokenized["input_ids"] if self.use_speech else None
            mask = input_tokenized["attention_mask"] if self.use_speech else None
            memory, mem_mask, B, E = eng.encode(video, ids, mask)
            if use_nucleus_sampling:
                if num_captions > 1:       # HF expands every input num_return_sequences times before sampling
                    d_ = memory.shape[1]
                    memory = memory.view(B, E, d_).repeat_interleave(num_captions, 0).reshape(B * num_captions * E, d_).contiguous()
                    mem_mask = mem_mask.repeat_interleave(num_captions, 0).contiguous()
                    B = B * num_captions
                seq = eng.generate_greedy(memory, mem_mask, B, E, max_new_tokens=max_length, repetition_penalty=repetition_penalty,
                                          min_length=min_length, sample=(float(top_p), float(temperature)))
            elif num_beams == 1:
                seq = eng.generate_greedy(memory, mem_mask, B, E, max_new_tokens=max_length,
                                          repetition_penalty=repetition_penalty, min_length=min_length)
            else:
                seq = eng.generate_beam(memory, mem_mask, B, E, num_beams=num_beams, max_new_tokens=max_length,
                                        length_penalty=length_penalty, repetition_penalty=repetition_penalty,
                                        min_length=min_length, num_return=num_captions)
        self.last_generated_ids = seq
        return self.t5_tokenizer.batch_decode(seq, skip_special_tokens=True)


def build_vid2seq_model(args, tokenizer):
    """Same factory as the reference's model/__init__.py:4-19."""
    return Vid2Seq(t5_path=args.model_name, num_features=args.max_feats, embed_dim=args.embedding_dim, depth=args.depth,
                   heads=args.heads, mlp_dim=args.mlp_dim, vis_drop=args.visual_encoder_dropout,
                   enc_drop=args.text_encoder_dropout, dec_drop=args.text_decoder_dropout, tokenizer=tokenizer,
                   num_bins=args.num_bins, label_smoothing=args.label_smoothing, use_speech=args.use_speech,
                   use_video=args.use_video)
