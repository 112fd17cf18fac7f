"""Host orchestration of the Vid2Seq train step over the op table (vidchapters_b200.ops.CudaOps).

This is the Python "host code calling hand-written sm_100a CUDA through a thin C-ABI extension" of BASELINE.json's
north_star: every arithmetic step below is one call into libvidchap.so; torch is used for device memory (torch.empty),
streams and integer/mask plumbing only.

Reference call graph being replaced (paths under /root/reference):
  model/vid2seq.py:58-98   Vid2Seq.forward            -> Vid2SeqEngine.forward
  model/vit.py:117-133     VisionTransformer.forward  -> _vit_fwd / _vit_bwd
  model/modeling_t5.py:930-1138 T5Stack.forward       -> _enc_fwd / _dec_fwd (+ _bwd)
  model/modeling_t5.py:1587-1738 lm_head + CE         -> _head_fwd / _head_bwd
  autograd backward of all of the above (dvc.py:113)  -> Vid2SeqEngine.backward
  dvc.py:114-126 clip / Adam / time-token renorm      -> Vid2SeqEngine.optimizer_step

Memory plan (HBM): parameters live in ONE fp32 buffer (`flat_p`) laid out by config.param_shapes, shadowed by a bf16
copy (`flat_pb`, what the GEMMs read through TMA), one fp32 gradient buffer (`flat_g`, accumulated into by atomics,
the single tensor the data-parallel all-reduce sends) and Adam m/v.  q,k,v weights are adjacent so the fused-QKV
[3*inner, d] matrix is a plain view.  The residual stream is fp32 [tokens, d]; every tensor that feeds a tensor-core
GEMM is stored bf16; activations needed by the backward are kept (4 GB at config 2 — no recompute on a 180 GB part).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .config import layout_order, param_shapes, vocab_size
from .ops import ACT_CE_GRAD, ACT_CE_STATS, ACT_GELU, ACT_GELU_BWD, ACT_NONE, ACT_RELU, ACT_RELU_BWD, NO_DROP, drop_spec

_ALIGN = 64  # elements; keeps every parameter 256 B (fp32) / 128 B (bf16) aligned for TMA


def relative_position_bucket(relative_position, bidirectional=True, num_buckets=32, max_distance=128):
    """Host-side bucket function, same torch ops as the reference (modeling_t5.py:397-443): fp32 log then truncation.
    Evaluated once per (Lq, Lk) into a LUT that the kernels index; the device never recomputes the log."""
    relative_buckets = torch.zeros_like(relative_position)
    if bidirectional:
        num_buckets //= 2
        relative_buckets = relative_buckets + (relative_position > 0).to(torch.long) * num_buckets
        relative_position = torch.abs(relative_position)
    else:
        relative_position = -torch.min(relative_position, torch.zeros_like(relative_position))
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    if_large = max_exact + (
        torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    if_large = torch.min(if_large, torch.full_like(if_large, num_buckets - 1))
    return relative_buckets + torch.where(is_small, relative_position, if_large)


def _restores_stream(fn):
    """forward / backward switch torch's current stream to the side stream for the visual-encoder chain.  If anything
    raises in between (a C-ABI status such as Lk > 1536, an out-of-memory error) the caller's stream is restored and
    joined with the side stream, so later torch work is not left running unsynchronised on the wrong stream."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        if self.device.type != "cuda":
            return fn(self, *a, **k)
        entry = torch.cuda.current_stream(self.device)
        try:
            return fn(self, *a, **k)
        except BaseException:
            torch.cuda.set_stream(entry)
            if self._side_stream is not None:
                try:
                    entry.wait_stream(self._side_stream)
                except Exception:
                    pass
            raise
    return wrapper


@dataclass
class _Sub:
    """One residual sub-layer's parameter names and static attributes."""
    kind: int            # 0 = T5 RMS norm (eps 1e-6), 1 = nn.LayerNorm (eps 1e-5)
    norm_w: str
    norm_b: Optional[str]
    H: int = 0
    scale: float = 1.0
    qkv_w: str = ""      # first of the adjacent q,k,v weights (or the fused ViT qkv)
    qkv_b: Optional[str] = None
    q_w: str = ""        # cross attention: separate q, fused k,v
    kv_w: str = ""
    o_w: str = ""
    o_b: Optional[str] = None
    w1: str = ""
    b1: Optional[str] = None
    w2: str = ""
    b2: Optional[str] = None
    act: int = ACT_RELU

    @property
    def eps(self):
        return 1e-5 if self.kind == 1 else 1e-6


class Vid2SeqEngine:
    @staticmethod
    def _fuse_cross_kv_default() -> bool:
        """Cross-attention K/V projections of all decoder layers as one GEMM (measured on B200, config 2: 31.10 ->
        30.59 ms/step, profiles/README.md round 2); VIDCHAP_FUSE_CROSS_KV=0 restores one GEMM pair per layer."""
        import os
        return os.environ.get("VIDCHAP_FUSE_CROSS_KV", "1") != "0"

    @staticmethod
    def _default_drop_seed() -> int:
        import os
        rank = 0
        try:
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()
            else:
                rank = int(os.environ.get("RANK", "0"))
        except Exception:
            rank = 0
        return (int(torch.initial_seed()) ^ (rank * 0x9E3779B1) ^ 0x5EED) & 0xFFFFFFFF

    @staticmethod
    def layout_of(cfg: dict, grouped_cross_kv: Optional[bool] = None):
        """name -> (offset, shape, numel) in the flat buffers.  grouped_cross_kv (default: VIDCHAP_FUSE_CROSS_KV) selects
        config.layout_order, which keeps the decoder's cross-attention k,v weights of all layers contiguous."""
        if grouped_cross_kv is None:
            grouped_cross_kv = Vid2SeqEngine._fuse_cross_kv_default()
        layout, off = {}, 0
        for name, shape in (layout_order(cfg) if grouped_cross_kv else param_shapes(cfg)):
            n = 1
            for s in shape:
                n *= s
            layout[name] = (off, tuple(shape), n)
            off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        return layout, off

    def __init__(self, cfg: dict, ops, device, label_smoothing: float = 0.1, use_video=True, use_speech=True,
                 flat_p: Optional[torch.Tensor] = None, fuse_cross_kv: Optional[bool] = None):
        self.cfg, self.ops, self.device = cfg, ops, torch.device(device)
        self.label_smoothing = label_smoothing
        self.use_video, self.use_speech = use_video, use_speech
        self.d, self.H, self.dff = cfg["d_model"], cfg["num_heads"], cfg["d_ff"]
        self.inner = self.H * cfg["d_kv"]
        assert cfg["d_kv"] == 64, "kernels are specialised for head_dim 64 (t5-base / t5-large / the ViT)"
        self.C, self.Hv, self.mlp = cfg["embed_dim"], cfg["heads"], cfg["mlp_dim"]
        assert self.C // self.Hv == 64
        self.V = vocab_size(cfg)
        # fuse_cross_kv (default on; VIDCHAP_FUSE_CROSS_KV=0 disables): the cross-attention K/V projections of all decoder
        # layers (same input: the encoder memory) as ONE GEMM [B*E, d] x [d, layers*2*inner] in the forward, and one
        # weight-gradient GEMM + one input-gradient GEMM (K = layers*2*inner) in the backward.  Needs the grouped
        # parameter layout, hence fixed at construction.
        self.fuse_cross_kv = self._fuse_cross_kv_default() if fuse_cross_kv is None else bool(fuse_cross_kv)
        self.layout, self.total = self.layout_of(cfg, self.fuse_cross_kv)
        dev = self.device
        if flat_p is not None:
            assert flat_p.numel() == self.total and flat_p.dtype == torch.float32 and flat_p.device == dev
            self.flat_p = flat_p
        else:
            self.flat_p = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.flat_pb = torch.zeros(self.total, dtype=torch.bfloat16, device=dev)
        # gradient buffer: 64 leading floats of slack in front of the parameters' gradients; the last of them carries the
        # step's loss through the data-parallel all-reduce of the first region (dvc.py:103's loss reduction rides along)
        self._g_store = torch.zeros(64 + self.total, dtype=torch.float32, device=dev)
        self.flat_g = self._g_store[64:]
        self.loss_slot = self._g_store[63:64]
        self.adam_m = None
        self.adam_v = None
        self.adam_step_count = 0
        self._scratch = torch.zeros(8, dtype=torch.float32, device=dev)
        self._luts: Dict[tuple, torch.Tensor] = {}
        self._one = torch.ones(1, dtype=torch.float32, device=dev)
        self.drop_rates = dict(vis=0.0, enc=0.0, dec=0.0)   # vis_drop / enc_drop / dec_drop of the reference ctor
        # user seed of the dropout streams: torch's seed (dvc.py:255-258 seeds `args.seed + rank` before the model is
        # built) mixed with the data-parallel rank, so that --seed changes the masks and replicas draw different ones;
        # every forward call then derives its own stream from (drop_seed, call counter)
        self.drop_seed = self._default_drop_seed()
        self._drop_calls = 0
        # The visual encoder (small 1600-row kernels that cannot fill the GPU) runs on a second stream next to the text
        # encoder, forward and backward; joined before the decoder / at the end of the backward.  Inside CUDA-graph capture
        # the fork/join become graph dependencies.  -3.7 % step time (profiles/README.md); VIDCHAP_DUAL_STREAM=0 disables.
        import os
        self.dual_stream = os.environ.get("VIDCHAP_DUAL_STREAM", "1") != "0" and self.device.type == "cuda"
        self.fused_ce = os.environ.get("VIDCHAP_FUSED_CE", "1") != "0"
        # decode steps: norm + linear fused skinny kernels (decode2.cu) instead of norm kernel + tcgen05 GEMM at M = batch
        self.decode_skinny = os.environ.get("VIDCHAP_DECODE_LINEAR", "1") != "0"
        self._side_stream = None
        self._build_specs()

    # ------------------------------------------------------------------ parameter views
    def p(self, name):
        o, shape, n = self.layout[name]
        return self.flat_p[o:o + n].view(shape)

    def g(self, name):
        o, shape, n = self.layout[name]
        return self.flat_g[o:o + n].view(shape)

    def pb(self, name, rows=None):
        """bf16 shadow as a 2-D matrix; rows= spans adjacent parameters (fused q,k,v)."""
        o, shape, n = self.layout[name]
        cols = shape[-1]
        r = rows if rows is not None else n // cols
        return self.flat_pb[o:o + r * cols].view(r, cols)

    def g2(self, name, rows=None):
        o, shape, n = self.layout[name]
        cols = shape[-1]
        r = rows if rows is not None else n // cols
        return self.flat_g[o:o + r * cols].view(r, cols)

    def gv(self, name):
        return None if name is None else self.g(name).view(-1)

    def pv(self, name):
        return None if name is None else self.p(name).view(-1)

    def sync_bf16(self):
        """Refresh the bf16 shadow from the fp32 masters (after load_state_dict / external optimiser steps)."""
        self.ops.cast_flat_bf16(self.flat_p, self.flat_pb)

    def _build_specs(self):
        cfg = self.cfg
        self.vit_blocks: List[tuple] = []
        for i in range(cfg["depth"]):
            p = f"visual_encoder.blocks.{i}."
            sa = _Sub(kind=1, norm_w=p + "norm1.weight", norm_b=p + "norm1.bias", H=self.Hv, scale=64 ** -0.5,
                      qkv_w=p + "attn.qkv.weight", qkv_b=p + "attn.qkv.bias", o_w=p + "attn.proj.weight",
                      o_b=p + "attn.proj.bias")
            ff = _Sub(kind=1, norm_w=p + "norm2.weight", norm_b=p + "norm2.bias", w1=p + "mlp.fc1.weight",
                      b1=p + "mlp.fc1.bias", w2=p + "mlp.fc2.weight", b2=p + "mlp.fc2.bias", act=ACT_GELU)
            self.vit_blocks.append((sa, ff))
        self.enc_blocks, self.dec_blocks = [], []
        for i in range(cfg["num_layers"]):
            p = f"t5_model.encoder.block.{i}.layer."
            sa = _Sub(kind=0, norm_w=p + "0.layer_norm.weight", norm_b=None, H=self.H, scale=1.0,
                      qkv_w=p + "0.SelfAttention.q.weight", o_w=p + "0.SelfAttention.o.weight")
            ff = _Sub(kind=0, norm_w=p + "1.layer_norm.weight", norm_b=None, w1=p + "1.DenseReluDense.wi.weight",
                      w2=p + "1.DenseReluDense.wo.weight", act=ACT_RELU)
            self.enc_blocks.append((sa, ff))
            p = f"t5_model.decoder.block.{i}.layer."
            sa = _Sub(kind=0, norm_w=p + "0.layer_norm.weight", norm_b=None, H=self.H, scale=1.0,
                      qkv_w=p + "0.SelfAttention.q.weight", o_w=p + "0.SelfAttention.o.weight")
            ca = _Sub(kind=0, norm_w=p + "1.layer_norm.weight", norm_b=None, H=self.H, scale=1.0,
                      q_w=p + "1.EncDecAttention.q.weight", kv_w=p + "1.EncDecAttention.k.weight",
                      o_w=p + "1.EncDecAttention.o.weight")
            ff = _Sub(kind=0, norm_w=p + "2.layer_norm.weight", norm_b=None, w1=p + "2.DenseReluDense.wi.weight",
                      w2=p + "2.DenseReluDense.wo.weight", act=ACT_RELU)
            self.dec_blocks.append((sa, ca, ff))
        self.enc_bias_name = "t5_model.encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"
        self.dec_bias_name = "t5_model.decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"

    def lut(self, Lq, Lk, bidirectional):
        key = (Lq, Lk, bidirectional)
        if key not in self._luts:
            rel = torch.arange(Lq + Lk - 1, dtype=torch.long) - (Lq - 1)
            self._luts[key] = relative_position_bucket(rel, bidirectional).to(torch.int32).to(self.device)
        return self._luts[key]

    # ------------------------------------------------------------------ dropout sites
    def _begin_dropout(self, training: bool):
        """Every forward call gets a fresh stream; every dropout site inside it a distinct seed.  Masks are regenerated
        in the backward from (seed, element index) — nothing is stored."""
        self._drop_on = bool(training) and any(v > 0 for v in self.drop_rates.values())
        self._drop_calls += 1
        self._drop_base = (self.drop_seed * 0x9E3779B1 + self._drop_calls * 0x7F4A7C15) & 0xFFFFFFFF
        self._drop_site = 0

    def _site(self, which: str):
        if not self._drop_on or self.drop_rates[which] <= 0:
            return NO_DROP
        self._drop_site += 1
        return drop_spec(self.drop_rates[which], (self._drop_base + self._drop_site * 0x632BE5AB) & 0xFFFFFFFF)

    # ------------------------------------------------------------------ second stream (VIDCHAP_DUAL_STREAM)
    def _fork(self):
        """Returns (main, side) with `side` ordered after everything enqueued on the current stream, or (None, None)."""
        if not self.dual_stream:
            return None, None
        if self._side_stream is None:
            # equal priorities by default: giving the visual encoder's (small) kernels a high-priority stream was
            # measured SLOWER on config 2 (29.4 vs 28.3 ms/step) although its backward tail is exposed at equal priority
            # — the text encoder's persistent GEMMs lose more than the tail costs.  VIDCHAP_SIDE_PRIORITY=1 to A/B.
            prio = -1 if os.environ.get("VIDCHAP_SIDE_PRIORITY", "0") == "1" else 0
            self._side_stream = torch.cuda.Stream(device=self.device, priority=prio)
        main = torch.cuda.current_stream(self.device)
        self._side_stream.wait_stream(main)
        return main, self._side_stream

    # ------------------------------------------------------------------ allocation helpers
    def _e(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def _z(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    _splits_cache: Dict[tuple, int] = {}

    def _splits(self, n_out, k_out, red):
        """Split-K factor of a weight-gradient GEMM dW[n_out, k_out] = dY^T.X over `red` tokens.  Mirrors the tile
        dispatch of vc_gemm_bf16 (csrc/gemm.cu: CTA-pair kernel, 256 x {256,128} tiles on 74 clusters) and minimises
        waves x (k-blocks per split + per-tile overhead): e.g. 768x768 over 4096 tokens runs as 8 splits = one full wave
        of 8 k-blocks instead of 13 splits = 1.6 waves of 5."""
        import os
        if os.environ.get("VIDCHAP_SPLITS_OLD") == "1":   # A/B switch for bench runs
            tiles = ((n_out + 127) // 128) * ((k_out + 255) // 256)
            return max(1, min((red + 63) // 64 // 4, -(-222 // tiles)))
        key = (n_out, k_out, red)
        if key in self._splits_cache:
            return self._splits_cache[key]
        kb = (red + 63) // 64
        clusters, overhead = 74, 6.0      # per-tile prologue + fp32 reduce-add epilogue, in units of one 256x256 k-block
        tiles1 = ((n_out + 127) // 128) * ((k_out + 255) // 256)
        s_base = max(1, min(kb // 4, -(-222 // tiles1)))   # baseline: ~1.5 waves of 128x256 tiles, >= 4 k-blocks per split
        costs = {}
        best = (float("inf"), 1)
        for s in range(1, max(1, min(16, kb // 4)) + 1):
            kbs = -(-kb // s)
            s_eff = -(-kb // kbs)
            mb2 = -(-n_out // 256)
            width = 128
            for cand in (256, 128):
                tiles = mb2 * (-(-k_out // cand)) * s_eff
                if tiles >= clusters or (cand == 128 and tiles >= clusters // 2):
                    width = cand
                    break
            else:
                tiles = mb2 * (-(-k_out // 128)) * s_eff
            waves = -(-tiles // clusters)
            cost = waves * (kbs * (width / 256.0) + overhead)
            costs[s] = cost
            if cost < best[0] - 1e-9:
                best = (cost, s_eff)
        # keep the baseline unless the model predicts a clear win (its constants are rough)
        pick = best[1] if best[0] < 0.85 * costs.get(min(s_base, max(costs)), float("inf")) else s_base
        self._splits_cache[key] = pick
        return pick

    # ------------------------------------------------------------------ sub-layers: forward
    def _sa_fwd(self, x0, sp: _Sub, B, L, bias_rel, kmask, causal, tape, dk="enc"):
        ops, M, D = self.ops, x0.shape[0], x0.shape[1]
        inner = sp.H * 64
        bf = torch.bfloat16
        h = self._e(M, D, dtype=bf)
        rstd = self._e(M)
        mean = self._e(M) if sp.kind == 1 else None
        ops.norm_fwd(sp.kind, x0, self.pv(sp.norm_w), self.pv(sp.norm_b), out_bf16=h, rstd=rstd, mean=mean, eps=sp.eps)
        qkv = self._e(M, 3 * inner, dtype=bf)
        ops.gemm(h, self.pb(sp.qkv_w, 3 * inner), qkv, bias=self.pv(sp.qkv_b))
        ctx = self._e(M, inner, dtype=bf)
        lse = self._e(B, sp.H, L)
        d_attn, d_out = self._site(dk), self._site(dk)
        # text encoder: rows past a sequence's last token are padding that no consumer can observe (masked keys get an
        # exactly-zero probability in every later attention) — whole query tiles of them are skipped, forward and backward
        qlk = dk == "enc" and kmask is not None
        ops.attn_fwd(qkv, qkv, qkv, q_col=0, k_col=inner, v_col=2 * inner, B=B, H=sp.H, Lq=L, Lk=L, out=ctx, lse2=lse,
                     bias_rel=bias_rel, kmask=kmask, causal=causal, scale=sp.scale, drop=d_attn, q_like_k=qlk)
        x1 = self._e(M, D)
        ops.gemm(ctx, self.pb(sp.o_w), x1, bias=self.pv(sp.o_b), residual=x0, drop=d_out)
        tape.append(dict(t="sa", sp=sp, x0=x0, h=h, rstd=rstd, mean=mean, qkv=qkv, ctx=ctx, lse=lse, B=B, L=L,
                         bias_rel=bias_rel, kmask=kmask, causal=causal, d_attn=d_attn, d_out=d_out, qlk=qlk))
        return x1

    def _ca_fwd(self, y1, sp: _Sub, B, S, memory, E, mem_mask, tape, kv=None):
        ops, M, D = self.ops, y1.shape[0], y1.shape[1]
        inner = sp.H * 64
        bf = torch.bfloat16
        h = self._e(M, D, dtype=bf)
        rstd = self._e(M)
        ops.norm_fwd(0, y1, self.pv(sp.norm_w), None, out_bf16=h, rstd=rstd, eps=sp.eps)
        qc = self._e(M, inner, dtype=bf)
        ops.gemm(h, self.pb(sp.q_w), qc)
        if kv is None:
            kv = self._e(B * E, 2 * inner, dtype=bf)
            ops.gemm(memory, self.pb(sp.kv_w, 2 * inner), kv)
        ctx = self._e(M, inner, dtype=bf)
        lse = self._e(B, sp.H, S)
        d_attn, d_out = self._site("dec"), self._site("dec")
        ops.attn_fwd(qc, kv, kv, q_col=0, k_col=0, v_col=inner, B=B, H=sp.H, Lq=S, Lk=E, out=ctx, lse2=lse,
                     bias_rel=None, kmask=mem_mask, causal=False, scale=1.0, drop=d_attn)
        y2 = self._e(M, D)
        ops.gemm(ctx, self.pb(sp.o_w), y2, residual=y1, drop=d_out)
        tape.append(dict(t="ca", sp=sp, x0=y1, h=h, rstd=rstd, qc=qc, kv=kv, ctx=ctx, lse=lse, B=B, S=S, E=E,
                         mem_mask=mem_mask, d_attn=d_attn, d_out=d_out))
        return y2

    def _ff_fwd(self, x1, sp: _Sub, tape, dk="enc"):
        ops, M, D = self.ops, x1.shape[0], x1.shape[1]
        bf = torch.bfloat16
        h = self._e(M, D, dtype=bf)
        rstd = self._e(M)
        mean = self._e(M) if sp.kind == 1 else None
        ops.norm_fwd(sp.kind, x1, self.pv(sp.norm_w), self.pv(sp.norm_b), out_bf16=h, rstd=rstd, mean=mean, eps=sp.eps)
        w1 = self.pb(sp.w1)
        act = self._e(M, w1.shape[0], dtype=bf)
        pre = self._e(M, w1.shape[0], dtype=bf) if sp.act == ACT_GELU else None
        d_act, d_out = self._site(dk), self._site(dk)
        ops.gemm(h, w1, act, bias=self.pv(sp.b1), act=sp.act, pre_out=pre, drop=d_act)
        x2 = self._e(M, D)
        ops.gemm(act, self.pb(sp.w2), x2, bias=self.pv(sp.b2), residual=x1, drop=d_out)
        tape.append(dict(t="ff", sp=sp, x0=x1, h=h, rstd=rstd, mean=mean, act=act, pre=pre, d_act=d_act, d_out=d_out))
        return x2

    # ------------------------------------------------------------------ sub-layers: backward
    def _wgrad(self, dy, x, gname, rows=None, tag=None):
        """g[N,K] += dy[M,N]^T @ x[M,K]  (reduction over the token dimension, split-K with fp32 atomics).
        (A variant that ran these GEMMs on their own stream, event-ordered against the scratch buffers, measured no gain
        on B200 — 31.35 vs 31.10 ms/step — and was removed; profiles/README.md round 2.)"""
        out = self.g2(gname, rows)
        splits = self._splits(out.shape[0], out.shape[1], dy.shape[0])
        self.ops.gemm(dy, x, out, a_mn=True, b_mn=True, atomic=True, splits=splits)

    def _ff_bwd(self, r, dx, dxb, ws, next_drop=NO_DROP):
        """dxb arrives already masked by this sub-layer's output dropout (r["d_out"]); the norm backward at the end
        emits the bf16 copy masked by `next_drop` = output dropout of the sub-layer processed next."""
        ops, sp = self.ops, r["sp"]
        M, D = dx.shape
        dff = r["act"].shape[1]
        if sp.b2:
            ops.colsum_bf16(dxb, self.gv(sp.b2))
        self._wgrad(dxb, r["act"], sp.w2, tag="dxb")
        dact = ws["dact"][:M * dff].view(M, dff)
        if sp.act == ACT_GELU:
            ops.gemm(dxb, self.pb(sp.w2), dact, b_mn=True, act=ACT_GELU_BWD, aux=r["pre"], drop=r["d_act"])
        else:
            ops.gemm(dxb, self.pb(sp.w2), dact, b_mn=True, act=ACT_RELU_BWD, aux=r["act"], drop=r["d_act"])
        if sp.b1:
            ops.colsum_bf16(dact, self.gv(sp.b1))
        self._wgrad(dact, r["h"], sp.w1, tag="dact")
        dh = ws["dhb"][:M * D].view(M, D)
        ops.gemm(dact, self.pb(sp.w1), dh, b_mn=True)
        ops.norm_bwd(sp.kind, dh, r["x0"], self.pv(sp.norm_w), r["rstd"], r["mean"], dx=dx, dx_bf16=dxb,
                     accumulate_dx=True, dw=self.gv(sp.norm_w), db=self.gv(sp.norm_b), dxb_drop=next_drop)

    def _sa_bwd(self, r, dx, dxb, ws, dbias_rel, lut, next_drop=NO_DROP):
        ops, sp = self.ops, r["sp"]
        M, D = dx.shape
        B, L, H = r["B"], r["L"], sp.H
        inner = H * 64
        if sp.o_b:
            ops.colsum_bf16(dxb, self.gv(sp.o_b))
        self._wgrad(dxb, r["ctx"], sp.o_w, tag="dxb")
        dctx = ws["dctx"][:M * inner].view(M, inner)
        ops.gemm(dxb, self.pb(sp.o_w), dctx, b_mn=True)
        dqkv = ws["dqkv"][:M * 3 * inner].view(M, 3 * inner)
        dq_acc = ws["dq_acc"][:M * inner].view(M, inner)   # cleared inside attn_bwd
        delta = ws["delta"][:B * H * L].view(B, H, L)
        qkv = r["qkv"]
        ops.attn_bwd(qkv, qkv, qkv, q_col=0, k_col=inner, v_col=2 * inner, B=B, H=H, Lq=L, Lk=L, out=r["ctx"],
                     lse2=r["lse"], bias_rel=r["bias_rel"], kmask=r["kmask"], causal=r["causal"], scale=sp.scale,
                     dout=dctx, do_col=0, delta=delta, dq_acc=dq_acc, dk=dqkv, dk_col=inner, dv=dqkv, dv_col=2 * inner,
                     dbias_rel=dbias_rel, bucket_lut=lut, drop=r["d_attn"], q_like_k=r["qlk"])
        ops.cast_f32_bf16(dq_acc, dqkv[:, :inner])
        if sp.qkv_b:
            ops.colsum_bf16(dqkv, self.gv(sp.qkv_b))
        self._wgrad(dqkv, r["h"], sp.qkv_w, rows=3 * inner, tag="dqkv")
        dh = ws["dhb"][:M * D].view(M, D)
        ops.gemm(dqkv, self.pb(sp.qkv_w, 3 * inner), dh, b_mn=True)
        ops.norm_bwd(sp.kind, dh, r["x0"], self.pv(sp.norm_w), r["rstd"], r["mean"], dx=dx, dx_bf16=dxb,
                     accumulate_dx=True, dw=self.gv(sp.norm_w), db=self.gv(sp.norm_b), dxb_drop=next_drop)

    def _ca_bwd(self, r, dx, dxb, ws, memory, dmem, next_drop=NO_DROP, dkv_out=None):
        """dkv_out: (fuse_cross_kv) this layer's slice of the all-layer dK/dV matrix; the K/V weight and memory gradients
        are then computed once for all layers by the caller."""
        ops, sp = self.ops, r["sp"]
        M, D = dx.shape
        B, S, E, H = r["B"], r["S"], r["E"], sp.H
        inner = H * 64
        self._wgrad(dxb, r["ctx"], sp.o_w, tag="dxb")
        dctx = ws["dctx"][:M * inner].view(M, inner)
        ops.gemm(dxb, self.pb(sp.o_w), dctx, b_mn=True)
        dq_acc = ws["dq_acc"][:M * inner].view(M, inner)   # cleared inside attn_bwd
        dq = ws["dqkv"][:M * inner].view(M, inner)
        dkv = dkv_out if dkv_out is not None else ws["dkv"][:B * E * 2 * inner].view(B * E, 2 * inner)
        delta = ws["delta"][:B * H * S].view(B, H, S)
        ops.attn_bwd(r["qc"], r["kv"], r["kv"], q_col=0, k_col=0, v_col=inner, B=B, H=H, Lq=S, Lk=E, out=r["ctx"],
                     lse2=r["lse"], bias_rel=None, kmask=r["mem_mask"], causal=False, scale=1.0, dout=dctx, do_col=0,
                     delta=delta, dq_acc=dq_acc, dk=dkv, dk_col=0, dv=dkv, dv_col=inner, dbias_rel=None, bucket_lut=None,
                     drop=r["d_attn"])
        ops.cast_f32_bf16(dq_acc, dq)
        self._wgrad(dq, r["h"], sp.q_w, tag="dqkv")
        dh = ws["dhb"][:M * D].view(M, D)
        ops.gemm(dq, self.pb(sp.q_w), dh, b_mn=True)
        ops.norm_bwd(0, dh, r["x0"], self.pv(sp.norm_w), r["rstd"], None, dx=dx, dx_bf16=dxb, accumulate_dx=True,
                     dw=self.gv(sp.norm_w), dxb_drop=next_drop)
        if dkv_out is not None:
            return
        self._wgrad(dkv, memory, sp.kv_w, rows=2 * inner, tag="dkv")
        ops.gemm(dkv, self.pb(sp.kv_w, 2 * inner), dmem, b_mn=True, residual=dmem)  # dmem += dkv @ Wkv

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _mask_u8(mask):
        m = mask if mask.dtype == torch.bool else (mask != 0)
        return m.contiguous().view(torch.uint8)

    @_restores_stream
    def forward(self, video, input_ids, input_mask, output_ids, output_mask, video_cached: bool = False,
                want_logits: bool = False, training: bool = False):
        """One Vid2Seq forward (vid2seq.py:58-98).  Returns (loss[1] device tensor, ctx) — ctx feeds backward().
        `video` is (B,T,768) features, or the cached projected output (B,T,d) when video_cached."""
        ops, cfg, d = self.ops, self.cfg, self.d
        bf = torch.bfloat16
        tape: List[dict] = []
        ctx = {"tape": tape}
        self._begin_dropout(training)
        B = output_ids.shape[0]
        T = video.shape[1] if self.use_video else 0
        L = input_ids.shape[1] if self.use_speech else 0
        S = output_ids.shape[1]
        E = T + L
        ctx.update(B=B, T=T, L=L, S=S, E=E, video_cached=video_cached)
        memory = self._e(B * E, d, dtype=bf)
        vid_f32 = None
        main_s, side_s = self._fork() if (self.use_video and self.use_speech and not video_cached) else (None, None)
        if side_s is not None:
            torch.cuda.set_stream(side_s)
        # ---------------- visual encoder (vit.py:117-133)
        if self.use_video:
            if video_cached:
                vid_f32 = video.reshape(B * T, d).contiguous().float()
                tmp = self._e(B * T, d, dtype=bf)
                ops.cast_f32_bf16(vid_f32, tmp)
                ops.copy_rows_bf16(tmp, memory, B, T, d, E, 0)
            else:
                C = self.C
                xv = self._e(B * T, C)
                d_pos = self._site("vis")
                ops.add_pos(video.contiguous().float(), self.p("visual_encoder.pos_embed"), xv.view(B, T, C),
                            cfg["num_features"], drop=d_pos)
                ctx["d_pos"] = d_pos
                for sa, ff in self.vit_blocks:
                    xv = self._sa_fwd(xv, sa, B, T, None, None, False, tape, dk="vis")
                    xv = self._ff_fwd(xv, ff, tape, dk="vis")
                rstd, mean = self._e(B * T), self._e(B * T)
                if d == 768 and C == 768:
                    vid_f32 = self._e(B * T, C)
                    # final LayerNorm lands directly in the decoder memory rows [b, 0:T)
                    ops.norm_fwd(1, xv, self.pv("visual_encoder.norm.weight"), self.pv("visual_encoder.norm.bias"),
                                 out_bf16=memory, rstd=rstd, mean=mean, eps=1e-5, rows_per_batch=T, out_batch_stride=E,
                                 out_row_offset=0)
                    ops.norm_fwd(1, xv, self.pv("visual_encoder.norm.weight"), self.pv("visual_encoder.norm.bias"),
                                 out_f32=vid_f32, eps=1e-5)
                    vn = None
                else:  # proj_v2t (vid2seq.py:54-56,64-65)
                    vn = self._e(B * T, C, dtype=bf)
                    ops.norm_fwd(1, xv, self.pv("visual_encoder.norm.weight"), self.pv("visual_encoder.norm.bias"),
                                 out_bf16=vn, rstd=rstd, mean=mean, eps=1e-5)
                    vid_f32 = self._e(B * T, d)
                    ops.gemm(vn, self.pb("proj_v2t.weight"), vid_f32, bias=self.pv("proj_v2t.bias"))
                    tmp = self._e(B * T, d, dtype=bf)
                    ops.cast_f32_bf16(vid_f32, tmp)
                    ops.copy_rows_bf16(tmp, memory, B, T, d, E, 0)
                ctx.update(vit_x=xv, vit_rstd=rstd, vit_mean=mean, vit_vn=vn)
        ctx["n_vit_tape"] = len(tape)
        if side_s is not None:
            torch.cuda.set_stream(main_s)      # the text encoder is enqueued on the main stream, concurrently
        # ---------------- text encoder (modeling_t5.py:930-1138)
        if self.use_speech:
            ids = input_ids.contiguous()
            x = self._e(B * L, d)
            d_emb_e = self._site("enc")
            ops.embed_fwd(ids, self.p("t5_model.shared.weight"), x, drop=d_emb_e)
            lut_e = self.lut(L, L, True)
            bias_e = self._e(self.H, 2 * L - 1)
            ops.bias_expand(self.p(self.enc_bias_name), lut_e, bias_e)
            kmask_e = self._mask_u8(input_mask)
            for sa, ff in self.enc_blocks:
                x = self._sa_fwd(x, sa, B, L, bias_e, kmask_e, False, tape, dk="enc")
                x = self._ff_fwd(x, ff, tape, dk="enc")
            rstd_e = self._e(B * L)
            d_fin_e = self._site("enc")
            ops.norm_fwd(0, x, self.pv("t5_model.encoder.final_layer_norm.weight"), None, out_bf16=memory, rstd=rstd_e,
                         eps=1e-6, rows_per_batch=L, out_batch_stride=E, out_row_offset=T, drop=d_fin_e)
            ctx.update(enc_x=x, enc_rstd=rstd_e, enc_ids=ids, lut_e=lut_e, d_emb_e=d_emb_e, d_fin_e=d_fin_e)
        ctx["n_enc_tape"] = len(tape)
        if side_s is not None:
            main_s.wait_stream(side_s)         # join: the decoder's cross-attention reads both halves of `memory`
        parts = []
        if self.use_video:
            parts.append(torch.ones(B, T, dtype=torch.uint8, device=self.device))
        if self.use_speech:
            parts.append(self._mask_u8(input_mask))
        mem_mask = torch.cat(parts, dim=1).contiguous()
        # ---------------- decoder
        out_ids = output_ids.contiguous()
        dec_in, labels = torch.empty_like(out_ids), torch.empty_like(out_ids)
        n_valid = self._e(1)
        ops.prepare_targets(out_ids, dec_in, labels, n_valid, 0)
        y = self._e(B * S, d)
        d_emb_d = self._site("dec")
        ops.embed_fwd(dec_in, self.p("t5_model.shared.weight"), y, drop=d_emb_d)
        lut_d = self.lut(S, S, False)
        bias_d = self._e(self.H, 2 * S - 1)
        ops.bias_expand(self.p(self.dec_bias_name), lut_d, bias_d)
        kmask_d = self._mask_u8(output_mask)
        kv_all = None
        if self.fuse_cross_kv:
            nl2 = len(self.dec_blocks) * 2 * self.inner
            kv_all = self._e(B * E, nl2, dtype=bf)
            ops.gemm(memory, self.pb(self.dec_blocks[0][1].kv_w, nl2), kv_all)
        for li, (sa, ca, ff) in enumerate(self.dec_blocks):
            y = self._sa_fwd(y, sa, B, S, bias_d, kmask_d, True, tape, dk="dec")
            kv_i = None if kv_all is None else kv_all[:, li * 2 * self.inner:(li + 1) * 2 * self.inner]
            y = self._ca_fwd(y, ca, B, S, memory, E, mem_mask, tape, kv=kv_i)
            y = self._ff_fwd(y, ff, tape, dk="dec")
        # ---------------- head: final norm * d^-0.5 (tied), lm_head, label-smoothed CE (modeling_t5.py:1709-1721)
        seq = self._e(B * S, d, dtype=bf)
        rstd_d = self._e(B * S)
        d_fin_d = self._site("dec")
        ops.norm_fwd(0, y, self.pv("t5_model.decoder.final_layer_norm.weight"), None, out_bf16=seq, rstd=rstd_d, eps=1e-6,
                     out_scale=d ** -0.5, drop=d_fin_d)
        ctx.update(d_emb_d=d_emb_d, d_fin_d=d_fin_d)
        Vp = (self.V + 7) // 8 * 8   # leading dimension padded to 16 B (vc.py's vocab 32100 is not a multiple of 8)
        loss = self._e(1)
        dlogits = self._e(B * S, Vp, dtype=bf)[:, :self.V]
        W = self.pb("t5_model.shared.weight")
        if want_logits or not self.fused_ce:
            # debug path / VIDCHAP_FUSED_CE=0: materialise the fp32 logits, then the stand-alone CE kernel
            logits = self._e(B * S, Vp)[:, :self.V]
            ops.gemm(seq, W, logits)
            ops.cross_entropy(logits, labels.view(-1), n_valid, self.label_smoothing, loss, dlogits)
        else:
            # LM head fused with the label-smoothed cross entropy (modeling_t5.py:1714-1721): the (B*S, V) fp32 logits
            # never exist.  GEMM 1 keeps only per-row partial (max, sum exp, sum z) per column block; a tiny kernel turns
            # them into lse + loss; GEMM 2 recomputes the logits tile by tile and writes d loss / d logits (bf16), the
            # operand of the two backward GEMMs, straight from its epilogue.
            logits = None
            tn = 256
            stats = self._e(B * S, 2 * ((self.V + tn - 1) // tn), 3)
            zy, lse = self._e(B * S), self._e(B * S)
            ce = dict(labels=labels.view(-1), n_valid=n_valid, smoothing=self.label_smoothing, stats=stats, zy=zy, lse=lse)
            ops.gemm(seq, W, None, act=ACT_CE_STATS, tile_n=tn, ce=ce)
            ops.ce_combine(stats, zy, labels.view(-1), n_valid, self.label_smoothing, self.V, lse, loss)
            ops.gemm(seq, W, dlogits, act=ACT_CE_GRAD, tile_n=tn, ce=ce)
        ctx.update(memory=memory, mem_mask=mem_mask, dec_in=dec_in, dec_x=y, dec_rstd=rstd_d, seq=seq, dlogits=dlogits,
                   lut_d=lut_d, vid_f32=vid_f32)
        if want_logits:
            ctx["logits"] = logits
        return loss, ctx

    # ------------------------------------------------------------------ backward
    def decoder_grad_range(self):
        """[lo, hi) of flat_g that is final once the head + decoder backward is done: every decoder parameter.
        (`shared` still receives the encoder's embedding gradient later.)"""
        names = [n for n in self.layout if n.startswith("t5_model.decoder.")]
        lo = min(self.layout[n][0] for n in names)
        hi = max(self.layout[n][0] + self.layout[n][2] for n in names)
        return lo, hi

    def _enc_groups(self):
        """Text-encoder layers in (up to) three groups, ascending; the backward walks them last to first.  Uneven on
        purpose: the group walked FIRST is the upper half (the visual encoder's backward runs next to it on the second
        stream and needs that long), the group walked LAST is as small as possible, because its gradient region — which
        also holds `shared`, final only after the embedding backward — is the one all-reduce that nothing overlaps."""
        nl = len(self.enc_blocks)
        if nl < 3:
            return [[i] for i in range(nl)]
        first = max(1, nl // 12)
        upper = nl // 2
        return [list(range(0, first)), list(range(first, nl - upper)), list(range(nl - upper, nl))]

    def dp_phases(self):
        """Data-parallel schedule of the backward: [(phase, [(lo, hi), ...])].  `backward(ctx, phase=p)` for p = 0, 1, ...
        in order is the whole backward; after phase p the listed ranges of `flat_g` are FINAL, so their all-reduce can
        run while the later phases compute (graphed.py).  Phase 0 = LM head + decoder; phase 1 = visual encoder (second
        stream) next to the last group of text-encoder layers; then the remaining groups; the last phase also folds the
        relative-position bias and the embedding gradient, which completes `shared` + the first group.  The ranges are
        ordered so that the LAST one to be reduced — the only one nothing overlaps — is as small as possible."""
        order = sorted(self.layout.items(), key=lambda kv: kv[1][0])
        first = lambda pred: next(off for n, (off, _, _) in order if pred(n))
        lo_dec, hi_dec = self.decoder_grad_range()
        enc_end = first(lambda n: n.startswith("t5_model.decoder."))
        dec_end = first(lambda n: n.startswith("visual_encoder."))
        assert enc_end == lo_dec and hi_dec <= dec_end
        if not (self.use_video and self.use_speech):
            return [(0, [(enc_end, dec_end)]), (1, [(0, enc_end), (dec_end, self.total)])]
        groups = self._enc_groups()
        start = [0] + [first(lambda n, g=g: n.startswith(f"t5_model.encoder.block.{g[0]}.")) for g in groups[1:]] + [enc_end]
        phases = [(0, [(enc_end, dec_end)])]
        for j in range(len(groups) - 1, -1, -1):
            regions = [(start[j], start[j + 1])]
            if j == len(groups) - 1:
                regions = [(dec_end, self.total)] + regions
            phases.append((len(groups) - j, regions))
        return phases

    @_restores_stream
    def backward(self, ctx, grad_loss: Optional[torch.Tensor] = None, grad_video: Optional[torch.Tensor] = None,
                 phase: Optional[int] = None):
        """Accumulates d(loss)/d(params) * grad_loss into flat_g.  Returns d/d(cached video) when the forward consumed
        a cached visual-encoder output (so it can flow back to the pass that produced it), else None.
        phase=None runs everything; phase=p runs one phase of `dp_phases()` (data-parallel overlap, graphed.py)."""
        if phase is not None and phase >= 1:
            if not (self.use_video and self.use_speech):
                return self._backward_rest(ctx, grad_video)
            groups = self._enc_groups()
            j = len(groups) - phase
            return self._backward_rest(ctx, grad_video, enc_layers=groups[j], do_vit=(phase == 1))
        ops, d = self.ops, self.d
        bf = torch.bfloat16
        tape = ctx["tape"]
        B, T, L, S, E = ctx["B"], ctx["T"], ctx["L"], ctx["S"], ctx["E"]
        gl = self._one if grad_loss is None else grad_loss.reshape(1).float().contiguous()
        Mmax = B * max(T, L, S)
        inner_max = max(self.inner, self.C)
        ws = dict(
            dact=self._e(Mmax * max(self.dff, self.mlp), dtype=bf), dh=self._e(Mmax * max(d, self.C)),
            dhb=self._e(Mmax * max(d, self.C), dtype=bf),   # d(normed activations): bf16 is enough, halves the traffic
            dctx=self._e(Mmax * inner_max, dtype=bf), dqkv=self._e(Mmax * 3 * inner_max, dtype=bf),
            dq_acc=self._e(Mmax * inner_max), delta=self._e(B * max(self.H, self.Hv) * max(T, L, S)),
            dkv=self._e(B * E * 2 * self.inner, dtype=bf))
        # ---- head
        seq, dlogits = ctx["seq"], ctx["dlogits"]
        ops.gemm(dlogits, seq, self.g2("t5_model.shared.weight"), a_mn=True, b_mn=True, atomic=True, alpha_dev=gl,
                 splits=self._splits(self.V, d, B * S))
        dseq = ws["dh"][:B * S * d].view(B * S, d)
        ops.gemm(dlogits, self.pb("t5_model.shared.weight"), dseq, b_mn=True, alpha_dev=gl)
        dy = self._e(B * S, d)
        dyb = self._e(B * S, d, dtype=bf)
        def out_drop(j):   # output dropout of tape record j (the sub-layer whose backward consumes the next bf16 dx)
            return tape[j]["d_out"] if j >= 0 else NO_DROP
        i = len(tape)
        ops.norm_bwd(0, dseq, ctx["dec_x"], self.pv("t5_model.decoder.final_layer_norm.weight"), ctx["dec_rstd"], None,
                     dx=dy, dx_bf16=dyb, accumulate_dx=False, dw=self.gv("t5_model.decoder.final_layer_norm.weight"),
                     scale=d ** -0.5, g_drop=ctx["d_fin_d"], dxb_drop=out_drop(i - 1))
        # ---- decoder blocks (reverse)
        nl, inner2 = len(self.dec_blocks), 2 * self.inner
        fuse = self.fuse_cross_kv
        dmem = self._e(B * E, d) if fuse else self._z(B * E, d)      # fused: written once by the all-layer GEMM below
        dkv_all = self._e(B * E, nl * inner2, dtype=bf) if fuse else None
        drel_d = self._z(self.H, 2 * S - 1)
        n_dec_end = ctx["n_enc_tape"]
        for li in reversed(range(nl)):
            self._ff_bwd(tape[i - 1], dy, dyb, ws, next_drop=out_drop(i - 2))
            self._ca_bwd(tape[i - 2], dy, dyb, ws, ctx["memory"], dmem, next_drop=out_drop(i - 3),
                         dkv_out=dkv_all[:, li * inner2:(li + 1) * inner2] if fuse else None)
            self._sa_bwd(tape[i - 3], dy, dyb, ws, drel_d, ctx["lut_d"],
                         next_drop=out_drop(i - 4) if i - 4 >= n_dec_end else NO_DROP)
            i -= 3
        assert i == n_dec_end
        if fuse:   # K/V projections of all layers at once: dW_kv = dKV^T . memory ; dmemory = dKV . W_kv
            kv0 = self.dec_blocks[0][1].kv_w
            self._wgrad(dkv_all, ctx["memory"], kv0, rows=nl * inner2)
            ops.gemm(dkv_all, self.pb(kv0, nl * inner2), dmem, b_mn=True)
        ops.bias_fold(drel_d, ctx["lut_d"], self.g(self.dec_bias_name))
        ops.embed_bwd(ctx["dec_in"].view(-1), dy, self.g("t5_model.shared.weight"), drop=ctx["d_emb_d"])
        ctx["_bwd_state"] = dict(dmem=dmem, ws=ws, dx=None, dxb=None, drel_e=None, vit_done=False, video_added=False,
                                 dvideo_out=None)
        if phase == 0:
            return None
        return self._backward_rest(ctx, grad_video)

    def _backward_rest(self, ctx, grad_video, enc_layers=None, do_vit=True):
        """Text encoder (the block indices `enc_layers`, walked downwards; None = all of them) and, if `do_vit`, the
        visual encoder — concurrently on the second stream when both run in this call."""
        ops, d = self.ops, self.d
        bf = torch.bfloat16
        tape = ctx["tape"]
        B, T, L, S, E = ctx["B"], ctx["T"], ctx["L"], ctx["S"], ctx["E"]
        st = ctx["_bwd_state"]
        dmem, ws = st["dmem"], st["ws"]
        n_vit, n_enc = ctx["n_vit_tape"], ctx["n_enc_tape"]
        nle = len(self.enc_blocks)
        if enc_layers is None:
            enc_layers = list(range(nle))
        enc_layers = sorted(enc_layers, reverse=True) if self.use_speech else []
        do_vit = do_vit and self.use_video and not st["vit_done"]

        def out_drop(j):
            return tape[j]["d_out"] if j >= 0 else NO_DROP
        if grad_video is not None and self.use_video and not st["video_added"]:
            dmem.view(B, E, d)[:, :T].add_(grad_video.reshape(B, T, d).to(dmem.dtype))
            st["video_added"] = True
        main_s, side_s = (None, None)
        ws_v = ws
        if do_vit and enc_layers and not ctx["video_cached"]:
            main_s, side_s = self._fork()      # forked HERE: the side stream must not wait for the text-encoder backward
            if side_s is not None:             # the two chains run concurrently: the visual one gets its own scratch
                Mv, Cv = B * T, self.C
                ws_v = dict(dact=self._e(Mv * self.mlp, dtype=bf), dh=self._e(Mv * max(d, Cv)),
                            dhb=self._e(Mv * max(d, Cv), dtype=bf), dctx=self._e(Mv * Cv, dtype=bf),
                            dqkv=self._e(Mv * 3 * Cv, dtype=bf), dq_acc=self._e(Mv * Cv), delta=self._e(B * self.Hv * T))
        # ---- text encoder
        if enc_layers:
            if st["dx"] is None:
                assert enc_layers[0] == nle - 1, "the text-encoder backward starts at its last layer"
                st["dx"] = self._e(B * L, d)
                st["dxb"] = self._e(B * L, d, dtype=bf)
                ops.norm_bwd(0, dmem, ctx["enc_x"], self.pv("t5_model.encoder.final_layer_norm.weight"), ctx["enc_rstd"],
                             None, dx=st["dx"], dx_bf16=st["dxb"], accumulate_dx=False,
                             dw=self.gv("t5_model.encoder.final_layer_norm.weight"), rows_per_batch=L, g_batch_stride=E,
                             g_row_offset=T, g_drop=ctx["d_fin_e"], dxb_drop=out_drop(n_enc - 1))
                st["drel_e"] = self._z(self.H, 2 * L - 1)
            dx, dxb, drel_e = st["dx"], st["dxb"], st["drel_e"]
            for li in enc_layers:
                i_sa, i_ff = n_vit + 2 * li, n_vit + 2 * li + 1
                self._ff_bwd(tape[i_ff], dx, dxb, ws, next_drop=out_drop(i_sa))
                self._sa_bwd(tape[i_sa], dx, dxb, ws, drel_e, ctx["lut_e"],
                             next_drop=out_drop(i_sa - 1) if i_sa - 1 >= n_vit else NO_DROP)
            if enc_layers[-1] == 0:
                ops.bias_fold(drel_e, ctx["lut_e"], self.g(self.enc_bias_name))
                ops.embed_bwd(ctx["enc_ids"].view(-1), dx, self.g("t5_model.shared.weight"), drop=ctx["d_emb_e"])
        # ---- visual encoder
        if do_vit:
            if side_s is not None:
                torch.cuda.set_stream(side_s)
            wsv = ws_v
            if ctx["video_cached"]:
                st["dvideo_out"] = dmem.view(B, E, d)[:, :T].contiguous()
            else:
                C = self.C
                i = n_vit
                dxv = self._e(B * T, C)
                dxvb = self._e(B * T, C, dtype=bf)
                if ctx["vit_vn"] is None:
                    ops.norm_bwd(1, dmem, ctx["vit_x"], self.pv("visual_encoder.norm.weight"), ctx["vit_rstd"],
                                 ctx["vit_mean"], dx=dxv, dx_bf16=dxvb, accumulate_dx=False,
                                 dw=self.gv("visual_encoder.norm.weight"), db=self.gv("visual_encoder.norm.bias"),
                                 rows_per_batch=T, g_batch_stride=E, g_row_offset=0, dxb_drop=out_drop(i - 1))
                else:
                    dvid = dmem.view(B, E, d)[:, :T].reshape(B * T, d).contiguous()
                    dvb = self._e(B * T, d, dtype=bf)
                    ops.cast_f32_bf16(dvid, dvb)
                    ops.colsum_bf16(dvb, self.gv("proj_v2t.bias"))
                    self._wgrad(dvb, ctx["vit_vn"], "proj_v2t.weight")
                    dvn = wsv["dh"][:B * T * C].view(B * T, C)
                    ops.gemm(dvb, self.pb("proj_v2t.weight"), dvn, b_mn=True)
                    ops.norm_bwd(1, dvn, ctx["vit_x"], self.pv("visual_encoder.norm.weight"), ctx["vit_rstd"],
                                 ctx["vit_mean"], dx=dxv, dx_bf16=dxvb, accumulate_dx=False,
                                 dw=self.gv("visual_encoder.norm.weight"), db=self.gv("visual_encoder.norm.bias"),
                                 dxb_drop=out_drop(i - 1))
                for _ in range(len(self.vit_blocks)):
                    self._ff_bwd(tape[i - 1], dxv, dxvb, wsv, next_drop=out_drop(i - 2))
                    self._sa_bwd(tape[i - 2], dxv, dxvb, wsv, None, None, next_drop=out_drop(i - 3))
                    i -= 2
                assert i == 0, i
                ops.add_pos_bwd(dxv, self.g("visual_encoder.pos_embed"), B, T, C, self.cfg["num_features"],
                                drop=ctx["d_pos"])
            st["vit_done"] = True
            if side_s is not None:
                torch.cuda.set_stream(main_s)
                main_s.wait_stream(side_s)
        finished = (st["vit_done"] or not self.use_video) and (not self.use_speech or (enc_layers and enc_layers[-1] == 0))
        if finished:
            ctx.pop("_bwd_state")
        return st["dvideo_out"]

    # ------------------------------------------------------------------ inference: encode + greedy decode
    @torch.no_grad()
    def encode(self, video, input_ids, input_mask):
        """Visual encoder + text encoder -> (memory bf16 [B*E, d], mem_mask u8 [B, E], B, E) as in the first half of
        Vid2Seq.generate (model/vid2seq.py:129-148).  Reuses the training forward's sub-layers with dropout off."""
        ops, cfg, d = self.ops, self.cfg, self.d
        bf = torch.bfloat16
        self._begin_dropout(False)
        tape: List[dict] = []
        B = video.shape[0] if self.use_video else input_ids.shape[0]
        T = video.shape[1] if self.use_video else 0
        L = input_ids.shape[1] if self.use_speech else 0
        E = T + L
        memory = self._e(B * E, d, dtype=bf)
        parts = []
        if self.use_video:
            C = self.C
            xv = self._e(B * T, C)
            ops.add_pos(video.contiguous().float(), self.p("visual_encoder.pos_embed"), xv.view(B, T, C), cfg["num_features"])
            for sa, ff in self.vit_blocks:
                xv = self._sa_fwd(xv, sa, B, T, None, None, False, tape, dk="vis")
                xv = self._ff_fwd(xv, ff, tape, dk="vis")
            if d == 768 and C == 768:
                ops.norm_fwd(1, xv, self.pv("visual_encoder.norm.weight"), self.pv("visual_encoder.norm.bias"),
                             out_bf16=memory, eps=1e-5, rows_per_batch=T, out_batch_stride=E, out_row_offset=0)
            else:
                vn = self._e(B * T, C, dtype=bf)
                ops.norm_fwd(1, xv, self.pv("visual_encoder.norm.weight"), self.pv("visual_encoder.norm.bias"),
                             out_bf16=vn, eps=1e-5)
                tmp = self._e(B * T, d, dtype=bf)
                ops.gemm(vn, self.pb("proj_v2t.weight"), tmp, bias=self.pv("proj_v2t.bias"))
                ops.copy_rows_bf16(tmp, memory, B, T, d, E, 0)
            parts.append(torch.ones(B, T, dtype=torch.uint8, device=self.device))
        if self.use_speech:
            x = self._e(B * L, d)
            ops.embed_fwd(input_ids.contiguous(), self.p("t5_model.shared.weight"), x)
            bias_e = self._e(self.H, 2 * L - 1)
            ops.bias_expand(self.p(self.enc_bias_name), self.lut(L, L, True), bias_e)
            kmask_e = self._mask_u8(input_mask)
            for sa, ff in self.enc_blocks:
                x = self._sa_fwd(x, sa, B, L, bias_e, kmask_e, False, tape, dk="enc")
                x = self._ff_fwd(x, ff, tape, dk="enc")
            ops.norm_fwd(0, x, self.pv("t5_model.encoder.final_layer_norm.weight"), None, out_bf16=memory, eps=1e-6,
                         rows_per_batch=L, out_batch_stride=E, out_row_offset=T)
            parts.append(kmask_e)
        tape.clear()
        return memory, torch.cat(parts, dim=1).contiguous(), B, E

    @torch.no_grad()
    def _decode_buffers(self, Bn):
        """Per-step activations of the incremental decoder for Bn sequences."""
        bf, d, inner = torch.bfloat16, self.d, self.inner
        Vp = (self.V + 7) // 8 * 8
        return dict(x=self._e(Bn, d), h=self._e(Bn, d, dtype=bf), q=self._e(Bn, inner, dtype=bf),
                    kvn=self._e(Bn, 2 * inner, dtype=bf), qkv=self._e(Bn, 3 * inner, dtype=bf),
                    ctxb=self._e(Bn, inner, dtype=bf), act=self._e(Bn, self.dff, dtype=bf),
                    logits=self._e(Bn, Vp)[:, :self.V])

    def _decode_step_logits(self, ids, buf, caches, kvmem, mem_mask, bias_d, pos, Bn, S, E, kv_div=0):
        """One incremental decoder step for Bn sequences (modeling_t5.py:484-525,551-556 with past_key_values): embeds
        `ids`, appends this position's K/V to `caches`, attends to positions <= *pos and to the encoder memory, and
        leaves the next-token logits in buf["logits"].  A fixed launch sequence (the position is a device scalar).
        Per layer: [RMS norm + fused q,k,v] -> cache append -> single-query attention -> [o + residual] -> [RMS norm + q]
        -> single-query cross-attention -> [o + residual] -> [RMS norm + wi + ReLU] -> [wo + residual]: nine launches,
        the bracketed ones vc_decode_linear (decode2.cu).  kv_div: beams of one video share its cross-attention K/V."""
        ops, d, H, inner = self.ops, self.d, self.H, self.inner
        x, h, q, kvn, qkv, ctxb, act, logits = (buf[k] for k in ("x", "h", "q", "kvn", "qkv", "ctxb", "act", "logits"))
        ops.embed_fwd(ids, self.p("t5_model.shared.weight"), x)
        skinny = self.decode_skinny
        for li, (sa, ca, ff) in enumerate(self.dec_blocks):
            if skinny:
                ops.decode_linear(x, self.pb(sa.qkv_w, 3 * inner), qkv, norm_w=self.pv(sa.norm_w), eps=1e-6)
                ops.kv_append(qkv[:, inner:], caches[li], pos)
                q_src = qkv
            else:
                ops.norm_fwd(0, x, self.pv(sa.norm_w), None, out_bf16=h, eps=1e-6)
                ops.gemm(h, self.pb(sa.qkv_w), q)                                   # q rows of the fused [q;k;v]
                k_name = sa.qkv_w.replace(".q.weight", ".k.weight")
                ops.gemm(h, self.pb(k_name, 2 * inner), kvn)                         # adjacent k,v weights
                ops.kv_append(kvn, caches[li], pos)
                q_src = q
            c2 = caches[li].view(Bn * S, 2 * inner)
            ops.attn_fwd(q_src, c2, c2, q_col=0, k_col=0, v_col=inner, B=Bn, H=H, Lq=1, Lk=S, out=ctxb, lse2=None,
                         bias_rel=bias_d, kmask=None, causal=True, scale=1.0, q_offset_dev=pos, kv_batch_rows=S,
                         bias_zero=S - 1, bias_len=2 * S - 1)
            if skinny:
                ops.decode_linear(ctxb, self.pb(sa.o_w), x, residual=x)
                ops.decode_linear(x, self.pb(ca.q_w), q, norm_w=self.pv(ca.norm_w), eps=1e-6)
            else:
                ops.gemm(ctxb, self.pb(sa.o_w), x, residual=x)
                ops.norm_fwd(0, x, self.pv(ca.norm_w), None, out_bf16=h, eps=1e-6)
                ops.gemm(h, self.pb(ca.q_w), q)
            ops.attn_fwd(q, kvmem[li], kvmem[li], q_col=0, k_col=0, v_col=inner, B=Bn, H=H, Lq=1, Lk=E, out=ctxb,
                         lse2=None, bias_rel=None, kmask=mem_mask, causal=False, scale=1.0, kv_batch_div=kv_div)
            if skinny:
                ops.decode_linear(ctxb, self.pb(ca.o_w), x, residual=x)
                ops.decode_linear(x, self.pb(ff.w1), act, norm_w=self.pv(ff.norm_w), eps=1e-6, relu=True)
                ops.decode_linear(act, self.pb(ff.w2), x, residual=x)
            else:
                ops.gemm(ctxb, self.pb(ca.o_w), x, residual=x)
                ops.norm_fwd(0, x, self.pv(ff.norm_w), None, out_bf16=h, eps=1e-6)
                ops.gemm(h, self.pb(ff.w1), act, act=ACT_RELU)
                ops.gemm(act, self.pb(ff.w2), x, residual=x)
        ops.norm_fwd(0, x, self.pv("t5_model.decoder.final_layer_norm.weight"), None, out_bf16=h, eps=1e-6,
                     out_scale=d ** -0.5)
        ops.gemm(h, self.pb("t5_model.shared.weight"), logits)
        return logits

    def _greedy_setup(self, memory, mem_mask, B, E, S, use_graph, select_on_device=True):
        """Cross-attention K/V of every layer (modeling_t5.py:516-524), self-attention caches, the decode state and ONE
        decode step captured as a CUDA graph (the step position is a device scalar, so the launch sequence is fixed)."""
        ops, H, inner = self.ops, self.H, self.inner
        bf, dev = torch.bfloat16, self.device
        nl = len(self.dec_blocks)
        if self.fuse_cross_kv:          # one GEMM for all layers (same input), per-layer column slices
            kv_all = self._e(B * E, nl * 2 * inner, dtype=bf)
            ops.gemm(memory, self.pb(self.dec_blocks[0][1].kv_w, nl * 2 * inner), kv_all)
            kvmem = [kv_all[:, li * 2 * inner:(li + 1) * 2 * inner] for li in range(nl)]
        else:
            kvmem = []
            for sa, ca, ff in self.dec_blocks:
                kv = self._e(B * E, 2 * inner, dtype=bf)
                ops.gemm(memory, self.pb(ca.kv_w, 2 * inner), kv)
                kvmem.append(kv)
        caches = [torch.zeros(B, S, 2 * inner, dtype=bf, device=dev) for _ in range(nl)]
        bias_d = self._e(H, 2 * S - 1)
        ops.bias_expand(self.p(self.dec_bias_name), self.lut(S, S, False), bias_d)
        st = dict(pos=torch.zeros(1, dtype=torch.int32, device=dev),
                  ids=torch.zeros(B, dtype=torch.int64, device=dev),            # decoder_start_token_id = 0
                  seq=torch.zeros(B, S + 1, dtype=torch.int64, device=dev),
                  done=torch.zeros(B, dtype=torch.uint8, device=dev), caches=caches, graph=None)
        buf = self._decode_buffers(B)

        st["logits"] = buf["logits"]

        def step():
            logits = self._decode_step_logits(st["ids"], buf, caches, kvmem, mem_mask, bias_d, st["pos"], B, S, E)
            if select_on_device:      # argmax + eos/pad bookkeeping + position advance inside the captured step
                ops.greedy_next(logits, st["done"], st["ids"], st["seq"], st["pos"], 1, 0)
                ops.step_advance(st["pos"])

        def rewind():
            st["pos"].zero_(); st["ids"].zero_(); st["seq"].zero_(); st["done"].zero_()
            for c in caches:
                c.zero_()

        st["step"], st["rewind"] = step, rewind
        if use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step()                                  # warm-up (attribute setup, LUT uploads) ...
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            rewind()                                    # ... then rewind the state
            # one private memory pool for every decode-step capture of this engine: a fresh pool per generate() call cost
            # ~60 cudaMalloc + ~75 cudaFree (44 ms of a 460 ms call at batch 64).  The previous call's graph is kept
            # alive until the new one exists so the pool is never released in between.
            if getattr(self, "_decode_pool", None) is None:
                self._decode_pool = torch.cuda.graph_pool_handle()
            st["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["graph"], pool=self._decode_pool):
                step()
            self._decode_graph_keepalive = st["graph"]
            rewind()
        return st

    @staticmethod
    def process_scores(scores, seq, cur_len, repetition_penalty, min_length, eos_id=1):
        """HF-4.28 RepetitionPenaltyLogitsProcessor + MinLengthLogitsProcessor on [n, V] scores given the token history
        seq[:, :cur_len] (start token included) — the two processors `Vid2Seq.generate`'s kwargs can switch on."""
        if repetition_penalty != 1.0:
            hist = seq[:, :cur_len]
            sc = scores.gather(1, hist)
            sc = torch.where(sc < 0, sc * repetition_penalty, sc / repetition_penalty)
            scores = scores.scatter(1, hist, sc)
        if cur_len < min_length:
            scores = scores.clone()
            scores[:, eos_id] = -float("inf")
        return scores

    def generate_greedy(self, memory, mem_mask, B, E, max_new_tokens=256, use_graph=None, check_every=16,
                        repetition_penalty=1.0, min_length=1, sample=None):
        """Greedy decoding with a KV cache (HF-4.28 greedy semantics: start id 0, argmax, sequences that emitted eos=1
        continue with pad=0, stop when all are done or after max_new_tokens).  Returns int64 [B, 1 + n] ids including
        the start token, like `t5_model.generate` (model/vid2seq.py:150-162 with num_beams=1).
        One decode step = a fixed launch sequence driven by a DEVICE step counter -> captured once as a CUDA graph."""
        S = int(max_new_tokens)
        if use_graph is None:
            use_graph = self.device.type == "cuda" and getattr(self.ops, "name", "") == "cuda"
        if repetition_penalty != 1.0 or min_length > 1 or sample is not None:
            return self._generate_host_select(memory, mem_mask, B, E, S, use_graph, repetition_penalty, min_length, sample)
        st = self._greedy_setup(memory, mem_mask, B, E, S, use_graph)
        n = 0
        while n < S:
            if n and n % check_every == 0 and bool(st["done"].all().item()):
                break
            if st["graph"] is not None:
                st["graph"].replay()
            else:
                st["step"]()
            n += 1
        return st["seq"][:, :n + 1].clone()

    def _generate_host_select(self, memory, mem_mask, B, E, S, use_graph, repetition_penalty, min_length, sample):
        """Greedy / nucleus-sampling decoding with HF's logits processors: the decode step (all the arithmetic up to the
        next-token logits) is the same captured launch sequence; the token choice — repetition penalty, min_length, then
        argmax or temperature + top-p + multinomial (`sample` = (top_p, temperature), torch's random stream) — is a few
        torch ops on the [B, V] logits per step."""
        ops = self.ops
        st = self._greedy_setup(memory, mem_mask, B, E, S, use_graph, select_on_device=False)
        seq, done = st["seq"], st["done"].bool()
        n = 0
        while n < S:
            if st["graph"] is not None:
                st["graph"].replay()
            else:
                st["step"]()
            scores = self.process_scores(st["logits"].float(), seq, n + 1, repetition_penalty, min_length)
            if sample is None:
                nxt = scores.argmax(-1)
            else:
                top_p, temperature = sample
                scores = scores / temperature
                srt, idx = torch.sort(scores, descending=False, dim=-1)
                remove = srt.softmax(-1).cumsum(-1) <= (1 - top_p)
                remove[..., -1:] = False
                scores = scores.masked_fill(remove.scatter(1, idx, remove), -float("inf"))
                nxt = torch.multinomial(scores.softmax(-1), 1).squeeze(1)
            nxt = torch.where(done, torch.zeros_like(nxt), nxt)
            done = done | (nxt == 1)
            seq[:, n + 1] = nxt
            st["ids"].copy_(nxt)
            ops.step_advance(st["pos"])
            n += 1
            if bool(done.all().item()):
                break
        return seq[:, :n + 1].clone()

    @torch.no_grad()
    def time_greedy_loop(self, memory, mem_mask, B, E, max_new_tokens=256):
        """Device time (ms, CUDA events) of `max_new_tokens` replays of the captured decode step — the decode loop alone,
        without encoders, cross K/V projection and capture (bench.py's HBM roofline of BASELINE configs[4])."""
        S = int(max_new_tokens)
        st = self._greedy_setup(memory, mem_mask, B, E, S, True)
        for _ in range(3):
            st["graph"].replay()
        st["rewind"]()
        torch.cuda.synchronize(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(S):
            st["graph"].replay()
        e1.record()
        torch.cuda.synchronize(self.device)
        return e0.elapsed_time(e1)

    def generate_beam(self, memory, mem_mask, B, E, num_beams=4, max_new_tokens=256, length_penalty=1.0, use_graph=None,
                      eos_id=1, pad_id=0, repetition_penalty=1.0, min_length=1, num_return=1):
        """Beam search with a KV cache (model/vid2seq.py:150-162 with num_beams > 1, i.e. the reference's default
        num_beams=4: HF-4.28 `beam_search` + `BeamSearchScorer`, early_stopping=False, length_penalty as given).
        Device: one decode step for all B*num_beams hypotheses (the greedy step's launch sequence), `vc_beam_topk`
        (log-softmax + beam scores + top 2*num_beams per batch item) and `vc_kv_reorder` (caches follow their beams,
        ping-pong between two cache sets so each parity's step is one fixed CUDA graph).  Host: the n-best bookkeeping on
        B x 2*num_beams candidates per step, as HF does.  Returns int64 [B, <= 1 + max_new_tokens] ids (start token,
        best hypothesis, eos, pad)."""
        ops, d, H, inner = self.ops, self.d, self.H, self.inner
        bf, dev = torch.bfloat16, self.device
        nb, S = int(num_beams), int(max_new_tokens)
        Bn, K2 = B * nb, 2 * nb
        if use_graph is None:
            use_graph = dev.type == "cuda" and getattr(ops, "name", "") == "cuda"
        # every beam of a batch item attends to the same memory (HF expands encoder outputs num_beams times): the K/V
        # projections are computed ONCE per video and the single-query attention kernel maps beam -> video (kv_batch_div)
        mask_x = mem_mask.repeat_interleave(nb, 0).contiguous()
        kvmem = []
        for sa, ca, ff in self.dec_blocks:
            kv = self._e(B * E, 2 * inner, dtype=bf)
            ops.gemm(memory, self.pb(ca.kv_w, 2 * inner), kv)
            kvmem.append(kv)
        nl = len(self.dec_blocks)
        caches = [[torch.zeros(Bn, S, 2 * inner, dtype=bf, device=dev) for _ in range(nl)] for _ in range(2)]
        bias_d = self._e(H, 2 * S - 1)
        ops.bias_expand(self.p(self.dec_bias_name), self.lut(S, S, False), bias_d)
        pos = torch.zeros(1, dtype=torch.int32, device=dev)
        ids = torch.zeros(Bn, dtype=torch.int64, device=dev)            # decoder_start_token_id = 0
        beam_scores = torch.zeros(Bn, dtype=torch.float32, device=dev)
        beam_idx = torch.zeros(Bn, dtype=torch.int32, device=dev)
        top_s = torch.zeros(B, K2, dtype=torch.float32, device=dev)
        top_t = torch.zeros(B, K2, dtype=torch.int32, device=dev)
        top_b = torch.zeros(B, K2, dtype=torch.int32, device=dev)
        buf = self._decode_buffers(Bn)

        processed = repetition_penalty != 1.0 or min_length > 1
        seq_dev = torch.zeros(Bn, S + 1, dtype=torch.int64, device=dev) if processed else None   # token history per beam

        def step(par):
            logits = self._decode_step_logits(ids, buf, caches[par], kvmem, mask_x, bias_d, pos, Bn, S, E, kv_div=nb)
            if not processed:     # log-softmax + beam scores + top 2*num_beams in one kernel
                ops.beam_topk(logits, beam_scores, nb, top_s, top_t, top_b)
            ops.step_advance(pos)

        def rewind():
            pos.zero_(); ids.zero_()
            init = torch.zeros(B, nb)
            init[:, 1:] = -1e9                      # HF: only beam 0 of every batch item is alive at the start
            beam_scores.copy_(init.view(-1))
            for par in range(2):
                for c in caches[par]:
                    c.zero_()

        graphs = [None, None]
        if use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step(0)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for par in range(2):
                rewind()
                graphs[par] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graphs[par]):
                    step(par)
        rewind()

        # ---- host-side n-best bookkeeping (HF-4.28 BeamSearchScorer.process / finalize, early_stopping=False)
        hyps = [[] for _ in range(B)]          # per batch item: list of (score, token list)
        worst = [1e9] * B
        done = [False] * B
        seqs = [[0] for _ in range(Bn)]        # tokens of every live beam (start token included)
        max_length = 1 + S
        par = 0

        def add_hyp(b, tokens, sum_logprobs):
            score = sum_logprobs / (len(tokens) ** length_penalty)    # length BEFORE the eos token, as HF
            if len(hyps[b]) < nb or score > worst[b]:
                hyps[b].append((score, list(tokens)))
                if len(hyps[b]) > nb:
                    srt = sorted((sc, i) for i, (sc, _) in enumerate(hyps[b]))
                    del hyps[b][srt[0][1]]
                    worst[b] = srt[1][0]
                else:
                    worst[b] = min(score, worst[b])

        while True:
            if graphs[par] is not None:
                graphs[par].replay()
            else:
                step(par)
            cur_len = len(seqs[0])
            if processed:
                # HF applies the logits processors to the LOG-PROBABILITIES in beam search, then adds the beam scores
                sc = self.process_scores(torch.log_softmax(buf["logits"].float(), -1), seq_dev, cur_len, repetition_penalty,
                                         min_length, eos_id) + beam_scores[:, None]
                V_ = sc.shape[1]
                s_, i_ = torch.topk(sc.view(B, nb * V_), K2, dim=1, largest=True, sorted=True)
                top_s.copy_(s_); top_t.copy_((i_ % V_).to(torch.int32)); top_b.copy_((i_ // V_).to(torch.int32))
            ts, tt, tb = top_s.cpu(), top_t.cpu(), top_b.cpu()      # one small D2H + sync per step
            n_scores = [0.0] * Bn
            n_tokens = [pad_id] * Bn
            n_index = [0] * Bn
            for b in range(B):
                if done[b]:
                    continue                   # finished items: scores 0, pad tokens, beam index 0 (as HF)
                k = 0
                for rank in range(K2):
                    tok, sc, bi = int(tt[b, rank]), float(ts[b, rank]), b * nb + int(tb[b, rank])
                    if tok == eos_id:
                        if rank >= nb:
                            continue
                        add_hyp(b, seqs[bi], sc)
                    else:
                        n_scores[b * nb + k], n_tokens[b * nb + k], n_index[b * nb + k] = sc, tok, bi
                        k += 1
                    if k == nb:
                        break
                if not done[b] and len(hyps[b]) >= nb:
                    done[b] = worst[b] >= float(ts[b].max()) / cur_len ** length_penalty
            seqs = [seqs[n_index[j]] + [n_tokens[j]] for j in range(Bn)]
            if all(done) or len(seqs[0]) >= max_length:
                break
            ids.copy_(torch.tensor(n_tokens, dtype=torch.int64), non_blocking=False)
            beam_scores.copy_(torch.tensor(n_scores, dtype=torch.float32))
            beam_idx.copy_(torch.tensor(n_index, dtype=torch.int32))
            if processed:
                seq_dev.copy_(seq_dev[beam_idx.long()])
                seq_dev[:, cur_len] = ids
            n_rows = cur_len                   # cache rows written so far: positions 0 .. cur_len-1
            for li in range(nl):
                ops.kv_reorder(caches[par][li], caches[1 - par][li], beam_idx, n_rows)
            par ^= 1
        final_scores = n_scores
        best = []
        for b in range(B):
            if not done[b]:
                for j in range(nb):
                    add_hyp(b, seqs[b * nb + j], final_scores[b * nb + j])
            ranked = sorted(hyps[b], key=lambda t: t[0])
            for j in range(num_return):            # num_return_sequences: the n best hypotheses, best first
                best.append(ranked[-1 - j][1])
        sent_max = min(max(len(t) for t in best) + 1, max_length)
        out = torch.full((B * num_return, sent_max), pad_id, dtype=torch.int64)
        for b, t in enumerate(best):
            out[b, :len(t)] = torch.tensor(t, dtype=torch.int64)
            if len(t) < sent_max:
                out[b, len(t)] = eos_id
        return out.to(dev)

    # ------------------------------------------------------------------ optimiser tail (dvc.py:114-126)
    def zero_grad(self):
        self.flat_g.zero_()

    def optimizer_step(self, lr, betas=(0.9, 0.999), eps=1e-8, clip_max_norm=1.0, grad_scale=1.0, renorm=True):
        ops = self.ops
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat_p)
            self.adam_v = torch.zeros_like(self.flat_p)
        self.adam_step_count += 1
        norm_sq = None
        if clip_max_norm > 0:
            norm_sq = self._scratch[0:1]
            norm_sq.zero_()
            ops.sumsq(self.flat_g, norm_sq)
        ops.adam_step(self.flat_p, self.flat_g, self.adam_m, self.adam_v, self.flat_pb, lr=lr, beta1=betas[0],
                      beta2=betas[1], eps=eps, step=self.adam_step_count, norm_sq=norm_sq, clip_max_norm=clip_max_norm,
                      grad_scale=grad_scale)
        if renorm and self.cfg["num_bins"]:
            w = self.p("t5_model.shared.weight")
            wb = self.pb("t5_model.shared.weight")
            for _ in range(2):  # shared and lm_head are the same tensor: the reference renormalises it twice
                ops.renorm_time_tokens(w, wb, self.cfg["num_bins"], self._scratch[2:4])
        return norm_sq
