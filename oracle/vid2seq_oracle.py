"""TEST INFRASTRUCTURE — CPU/torch restatement of the reference Vid2Seq train step.

This is the ORACLE (checker) for the CUDA path in vidchapters_b200/.  It is a
plain-PyTorch functional restatement of the reference algorithm over a state
dict that uses the reference's own key names.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product package never does.

PINNING: tests/test_oracle_cpu.py checks this file against the real,
unmodified reference (imported through oracle/ref_shim.py) when /root/reference
is present, and tests/golden/*.pt hold vectors minted from the real reference
by oracle/make_golden.py, against which this file is checked everywhere.
The reference itself ships no tests/golden vectors for this path,
so parity is pinned to the reference's own outputs, not to reference tests.
beam_search_decode() restates transformers==4.28.0's beam search (third-party,
absent from /root/reference, 4.28 itself not installable offline).  It is pinned
to the nearest real implementation that runs here: stock HuggingFace `generate`
of transformers 5.5 on a small trained T5 decoder (oracle/make_golden_beam.py ->
tests/golden/beam_hf.pt, tests/test_oracle_cpu.py) — exact on every case at the
reference's settings, with the two 4.28->5.5 differences found stated in that test
(fill value after eos; still-running beams when max_length is reached).

Each function cites the reference lines it restates (paths under /root/reference).

Two arithmetic modes:
  emulate_bf16=False : fp32 everywhere == the reference's numerics (tier B).
  emulate_bf16=True  : every matmul operand (Linear input+weight, Q.K^T, P.V and
                       their backward counterparts) is rounded to bf16 and
                       accumulated in fp32 — what a bf16 tensor-core kernel with
                       fp32 accumulation computes (tier A).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

NEG_MIN = torch.finfo(torch.float32).min  # HF additive mask constant, modeling_t5.py:996,1005 (third-party helper)


def _r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _EmuMatmul(torch.autograd.Function):
    """C = A @ B with operands rounded to bf16 (forward AND backward), fp32 accumulate.
    acc64=True accumulates the (exact) bf16 x bf16 products in fp64 and rounds the sum once to fp32: the
    accumulation-order-free version of the same arithmetic.  tests/ use the distance between the two as the measured
    noise floor of tier A (what ANY fp32-accumulating bf16 kernel is allowed to differ by)."""

    @staticmethod
    def forward(ctx, a, b, acc64=False):
        ar, br = _r(a), _r(b)
        ctx.save_for_backward(ar, br)
        ctx.acc64 = acc64
        if acc64:
            return (ar.double() @ br.double()).float()
        return ar @ br

    @staticmethod
    def backward(ctx, g):
        ar, br = ctx.saved_tensors
        gr = _r(g)
        if ctx.acc64:
            gd = gr.double()
            return (gd @ br.double().transpose(-1, -2)).float(), (ar.double().transpose(-1, -2) @ gd).float(), None
        return gr @ br.transpose(-1, -2), ar.transpose(-1, -2) @ gr, None


class DropPlan:
    """Replays the engine's counter-based dropout (vidchapters_b200/engine.py::_site, csrc/ptx.cuh::drop_keep): the
    k-th dropout site with a positive rate gets seed base + k*0x632BE5AB; element index = flat index of the tensor."""

    def __init__(self, rates: dict, base: int):
        self.rates, self.base, self.k = rates, base, 0

    def mask(self, which: str, x: torch.Tensor):
        from oracle.torch_ops import _idx, drop_mask
        p = self.rates.get(which, 0.0)
        if p <= 0:
            return None
        self.k += 1
        p16 = int(round(p * 65536.0))
        return drop_mask(((self.base + self.k * 0x632BE5AB) & 0xFFFFFFFF, p16), _idx(x.shape, x.device))


class Arith:
    def __init__(self, emulate_bf16: bool, flash_rounding: bool = False, drop_plan: "DropPlan" = None,
                 acc64: bool = False, trace: Optional[list] = None, torch_dropout: Optional[dict] = None):
        self.emu = emulate_bf16
        self.dp = drop_plan
        # torch_dropout {"vis","enc","dec": p}: nn.Dropout / F.dropout at every reference site with torch's own random
        # stream — what the reference itself runs in training mode (used by bench.py's CPU arm, not by parity tests)
        self.tdrop = torch_dropout
        self.acc64 = acc64 and emulate_bf16
        # trace: list that receives (name, x_in, x_out) for every residual sub-layer (detached) — the teacher-forced
        # per-sub-layer parity test feeds x_in to the CUDA sub-layer and compares with x_out
        self.trace = trace
        # flash_rounding: the probabilities that enter P.V are the UN-normalised exp(s - rowmax) rounded to bf16, the
        # division by the row sum happens after the product — the rounding points of any flash-attention kernel with
        # bf16 operands.  (Normalised-then-rounded P, the default, is what eager bf16 attention does; the two differ by
        # ~3.5e-3 rel-L2 on the logits purely through the rounding realisation, see tests/test_e2e_gpu.py.)
        self.flash = flash_rounding and emulate_bf16

    def dropout(self, which, x):
        """nn.Dropout / F.dropout site (training mode) with the replayed mask; identity without a plan."""
        if self.dp is None:
            if self.tdrop and self.tdrop.get(which, 0.0) > 0:
                return F.dropout(x, self.tdrop[which], training=True)
            return x
        m = self.dp.mask(which, x)
        return x if m is None else x * m

    def softmax_pv(self, scores, v, which=None):
        """softmax -> dropout on the probabilities (modeling_t5.py:569-574, vit.py:48-49) -> . V"""
        if not self.flash:
            return self.matmul(self.dropout(which, F.softmax(scores.float(), dim=-1)), v)
        # kernel rounding points: P~ = 2^(s*log2e - ceil(rowmax*log2e)) is rounded to bf16 (the integer exponent offset
        # makes the rounding independent of the kernel's tiling), dropped, multiplied by V, then divided by sum(P~).
        s2 = scores * 1.4426950408889634
        e = torch.exp2(s2 - torch.ceil(s2.max(-1, keepdim=True).values))
        m = self.dp.mask(which, e) if self.dp is not None else None
        if m is None:
            return self.matmul(e, v) / e.sum(-1, keepdim=True)
        # the kernel rounds the KEPT probabilities unscaled and folds 1/(1-p) into the final normalisation
        return self.matmul(e * (m > 0).to(e.dtype), v) / e.sum(-1, keepdim=True) * m.max()

    def matmul(self, a, b):
        if self.emu:
            return _EmuMatmul.apply(a, b, self.acc64)
        return a @ b

    def rec(self, name, x_in, x_out):
        if self.trace is not None:
            self.trace.append((name, x_in.detach(), x_out.detach()))
        return x_out

    def linear(self, x, w, b=None):
        y = self.matmul(x, w.t())
        return y if b is None else y + b


# --------------------------------------------------------------------------------------
# model/vit.py
# --------------------------------------------------------------------------------------
def vit_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, ar: Arith, pfx="visual_encoder.") -> torch.Tensor:
    """model/vit.py:117-133 (VisionTransformer.forward), :73-76 (Block), :38-55 (Attention), :16-22 (Mlp)."""
    B, N, C = x.shape
    H = cfg["heads"]
    pos = sd[pfx + "pos_embed"]
    if N != pos.shape[1]:  # vit.py:119-123 nearest interpolation of the time embedding
        pos = F.interpolate(pos.transpose(1, 2), size=N, mode="nearest").transpose(1, 2)
    x = ar.dropout("vis", x + pos)  # pos_drop, vit.py:126
    for i in range(cfg["depth"]):
        p = f"{pfx}blocks.{i}."
        h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        qkv = ar.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        qkv = qkv.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ar.matmul(q, k.transpose(-2, -1)) * ((C // H) ** -0.5)  # vit.py:47
        o = ar.softmax_pv(attn, v, "vis").transpose(1, 2).reshape(B, N, C)        # vit.py:48-51
        x = ar.rec(f"vit.{i}.sa", x, x + ar.dropout("vis", ar.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])))  # proj_drop
        h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        h = ar.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = ar.dropout("vis", F.gelu(h))  # nn.GELU() exact erf + drop, vit.py:9,19-20
        x = ar.rec(f"vit.{i}.ff", x, x + ar.dropout("vis", ar.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])))  # vit.py:21-22
    return ar.rec("vit.final", x, F.layer_norm(x, (C,), sd[pfx + "norm.weight"], sd[pfx + "norm.bias"], 1e-5))


# --------------------------------------------------------------------------------------
# model/modeling_t5.py
# --------------------------------------------------------------------------------------
def t5_layer_norm(x, w, eps=1e-6):
    """modeling_t5.py:254-277 (T5LayerNorm): RMS norm, fp32 statistics, no mean, no bias."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def relative_position_bucket(relative_position, bidirectional=True, num_buckets=32, max_distance=128):
    """modeling_t5.py:397-443.  Integer bucket of (memory_pos - query_pos); fp32 log then truncation."""
    relative_buckets = torch.zeros_like(relative_position)
    if bidirectional:
        num_buckets //= 2
        relative_buckets = relative_buckets + (relative_position > 0).to(torch.long) * num_buckets
        relative_position = torch.abs(relative_position)
    else:
        relative_position = -torch.min(relative_position, torch.zeros_like(relative_position))
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    large = max_exact + (
        torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return relative_buckets + torch.where(is_small, relative_position, large)


def compute_bias(table: torch.Tensor, qlen: int, klen: int, bidirectional: bool) -> torch.Tensor:
    """modeling_t5.py:445-460 -> (1,H,q,k)."""
    ctx = torch.arange(qlen, dtype=torch.long, device=table.device)[:, None]
    mem = torch.arange(klen, dtype=torch.long, device=table.device)[None, :]
    bucket = relative_position_bucket(mem - ctx, bidirectional=bidirectional)
    return table[bucket].permute(2, 0, 1).unsqueeze(0)


def t5_attention(sd, p, H, dkv, x, kv, position_bias, ar: Arith, which=None):
    """modeling_t5.py:462-588 (no cache, no head mask, no dropout): UNSCALED q.k^T + bias, fp32 softmax."""
    B, Lq, _ = x.shape
    Lk = kv.shape[1]
    q = ar.linear(x, sd[p + "q.weight"]).view(B, Lq, H, dkv).transpose(1, 2)
    k = ar.linear(kv, sd[p + "k.weight"]).view(B, Lk, H, dkv).transpose(1, 2)
    v = ar.linear(kv, sd[p + "v.weight"]).view(B, Lk, H, dkv).transpose(1, 2)
    scores = ar.matmul(q, k.transpose(3, 2))
    scores = scores + position_bias
    o = ar.softmax_pv(scores, v, which).transpose(1, 2).contiguous().view(B, Lq, H * dkv)  # fp32 softmax, :569-580
    return ar.linear(o, sd[p + "o.weight"])


def t5_ff(sd, p, x, ar: Arith, which=None):
    """modeling_t5.py:339-354 + :296-311 (T5LayerFF / T5DenseActDense, ReLU)."""
    h = t5_layer_norm(x, sd[p + "layer_norm.weight"])
    h = ar.linear(h, sd[p + "DenseReluDense.wi.weight"])
    h = ar.dropout(which, torch.relu(h))  # :306-307
    return ar.rec(p + "ff", x, x + ar.dropout(which, ar.linear(h, sd[p + "DenseReluDense.wo.weight"])))  # :353


def t5_encoder(sd, cfg, embeds, mask, ar: Arith, pfx="t5_model.encoder."):
    """modeling_t5.py:930-1138 (T5Stack.forward, encoder), :659-769 (T5Block)."""
    H, dkv = cfg["num_heads"], cfg["d_kv"]
    B, L, _ = embeds.shape
    ext = (1.0 - mask[:, None, None, :].to(torch.float32)) * NEG_MIN  # get_extended_attention_mask (HF)
    table = sd[pfx + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    bias = compute_bias(table, L, L, True) + ext  # modeling_t5.py:543-559 (layer 0 builds, others reuse :1092-1097)
    x = ar.dropout("enc", embeds)  # :1019
    for i in range(cfg["num_layers"]):
        p = f"{pfx}block.{i}."
        h = t5_layer_norm(x, sd[p + "layer.0.layer_norm.weight"])
        x = ar.rec(p + "layer.0.sa", x,
                   x + ar.dropout("enc", t5_attention(sd, p + "layer.0.SelfAttention.", H, dkv, h, h, bias, ar, "enc")))  # :618
        x = t5_ff(sd, p + "layer.1.", x, ar, "enc")
    return ar.rec("enc.final", x, ar.dropout("enc", t5_layer_norm(x, sd[pfx + "final_layer_norm.weight"])))  # :1113-1114


def shift_right(labels):
    """modeling_t5.py:845-868: prepend decoder_start_token_id=0, -100 -> pad 0."""
    s = labels.new_zeros(labels.shape)
    s[..., 1:] = labels[..., :-1].clone()
    s[..., 0] = 0
    return s.masked_fill(s == -100, 0)


def t5_decoder(sd, cfg, dec_ids, dec_mask, enc_h, enc_mask, ar: Arith, pfx="t5_model.decoder."):
    """modeling_t5.py:930-1138 (decoder stack): causal self-attn (+rel bias), cross-attn (zero bias), FF."""
    H, dkv = cfg["num_heads"], cfg["d_kv"]
    B, S = dec_ids.shape
    x = ar.dropout("dec", sd["t5_model.shared.weight"][dec_ids])  # modeling_t5.py:972,1019
    causal = torch.tril(torch.ones(S, S, device=x.device))[None, :, :] * dec_mask[:, None, :].to(torch.float32)
    ext = (1.0 - causal[:, None, :, :]) * NEG_MIN  # create_extended_attention_mask_for_decoder (HF)
    table = sd[pfx + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    self_bias = compute_bias(table, S, S, False) + ext
    cross_bias = (1.0 - enc_mask[:, None, None, :].to(torch.float32)) * NEG_MIN  # invert_attention_mask + zeros :544-547
    for i in range(cfg["num_layers"]):
        p = f"{pfx}block.{i}."
        h = t5_layer_norm(x, sd[p + "layer.0.layer_norm.weight"])
        x = ar.rec(p + "layer.0.sa", x,
                   x + ar.dropout("dec", t5_attention(sd, p + "layer.0.SelfAttention.", H, dkv, h, h, self_bias, ar, "dec")))
        h = t5_layer_norm(x, sd[p + "layer.1.layer_norm.weight"])
        x = ar.rec(p + "layer.1.ca", x,
                   x + ar.dropout("dec", t5_attention(sd, p + "layer.1.EncDecAttention.", H, dkv, h, enc_h, cross_bias, ar, "dec")))
        x = t5_ff(sd, p + "layer.2.", x, ar, "dec")
    return ar.rec("dec.final", x, ar.dropout("dec", t5_layer_norm(x, sd[pfx + "final_layer_norm.weight"])))


def vid2seq_forward(sd, cfg, video, input_ids, input_mask, output_ids, output_mask, *, emulate_bf16=False,
                    label_smoothing=0.1, video_is_cached=False, flash_rounding=False, drop_plan=None, use_video=True,
                    use_speech=True, acc64=False, trace=None, torch_dropout=None):
    """model/vid2seq.py:58-98 (+ modeling_t5.py:1587-1738).  Dropout-free (p=0 / eval) restatement.
    use_video / use_speech = the reference's --no_video / --no_speech variants (vid2seq.py:59-84): the decoder's memory
    is the visual tokens, the text-encoder states, or their concatenation.

    Returns dict(loss, logits (B,S,V), video (B,T,d), memory (B,T+L,d)).
    """
    ar = Arith(emulate_bf16, flash_rounding, drop_plan, acc64=acc64, trace=trace, torch_dropout=torch_dropout)
    d = cfg["d_model"]
    vid = None
    if use_video:
        if video_is_cached:
            vid = video
        else:
            vid = vit_forward(sd, cfg, video, ar)
            if d != 768:  # vid2seq.py:54-56,64-65
                vid = ar.linear(vid, sd["proj_v2t.weight"], sd["proj_v2t.bias"])
        atts_vis = torch.ones(vid.shape[:2], dtype=torch.long, device=vid.device)
    if use_speech:
        text = sd["t5_model.shared.weight"][input_ids]  # vid2seq.py:71
        enc = t5_encoder(sd, cfg, text, input_mask, ar)
    if use_video and use_speech:
        memory = torch.cat([vid, enc], dim=1)  # vid2seq.py:78
        mem_mask = torch.cat([atts_vis, input_mask.to(torch.long)], dim=1)
    elif use_video:
        memory, mem_mask = vid, atts_vis       # vid2seq.py:80-82
    else:
        memory, mem_mask = enc, input_mask.to(torch.long)   # vid2seq.py:83-84
    targets = output_ids.masked_fill(output_ids == 0, -100)  # vid2seq.py:86-88
    dec_in = shift_right(targets)
    seq = t5_decoder(sd, cfg, dec_in, output_mask, memory, mem_mask, ar)
    if trace is not None:
        trace.append(("memory", memory.detach(), mem_mask.detach()))
    seq = seq * (d ** -0.5)  # modeling_t5.py:1709-1712 (tied embeddings)
    logits = ar.linear(seq, sd["t5_model.shared.weight"])  # lm_head tied to shared, F9
    loss = F.cross_entropy(logits.view(-1, logits.size(-1)), targets.view(-1), ignore_index=-100,
                           label_smoothing=label_smoothing)  # modeling_t5.py:1721
    return {"loss": loss, "logits": logits, "video": vid, "memory": memory}


# --------------------------------------------------------------------------------------
# dvc.py step tail
# --------------------------------------------------------------------------------------
def clip_adam_renorm_(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: dict, *, lr=3e-4,
                      betas=(0.9, 0.999), eps=1e-8, clip_max_norm=1.0, num_bins=100):
    """dvc.py:112-126: clip_grad_norm_ (global L2), torch.optim.Adam (wd 0), time-token renorm.  In place."""
    names = list(params.keys())
    total = torch.sqrt(sum((grads[n].double() ** 2).sum() for n in names)).float()
    coef = 1.0
    if clip_max_norm > 0:
        coef = torch.clamp(clip_max_norm / (total + 1e-6), max=1.0)  # torch clip_grad_norm_
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    b1, b2 = betas
    for n in names:
        g = grads[n] * coef
        m = state.setdefault("m." + n, torch.zeros_like(params[n]))
        v = state.setdefault("v." + n, torch.zeros_like(params[n]))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        params[n].addcdiv_(m, denom, value=-lr / bc1)
    if num_bins:
        w = params["t5_model.shared.weight"]  # lm_head is the same tensor (tied): renorm applies twice, dvc.py:118-126
        for _ in range(2):
            frozen = torch.norm(w[:-num_bins], dim=1).mean(0)
            w[-num_bins:].div_(torch.norm(w[-num_bins:], dim=1).mean(0) / frozen)
    return total


def process_scores(scores, ids, repetition_penalty=1.0, min_length=1, eos_id=1):
    """The two HF logits processors `Vid2Seq.generate` can switch on (vid2seq.py:150-162 -> transformers 4.28
    RepetitionPenaltyLogitsProcessor, MinLengthLogitsProcessor): tokens already in `ids` (start token included) get
    score*p if negative else score/p; eos is forbidden while len(ids) < min_length.  `scores` are raw logits in greedy /
    sampling and log-probabilities in beam search, as in HF."""
    if repetition_penalty != 1.0:
        sc = scores.gather(1, ids)
        sc = torch.where(sc < 0, sc * repetition_penalty, sc / repetition_penalty)
        scores = scores.scatter(1, ids, sc)
    if ids.shape[1] < min_length:
        scores = scores.clone()
        scores[:, eos_id] = -float("inf")
    return scores


def top_p_filter(scores, top_p, temperature=1.0):
    """HF TemperatureLogitsWarper + TopPLogitsWarper (min_tokens_to_keep=1): the smallest set of tokens whose probability
    mass reaches top_p keeps its logits, everything else becomes -inf."""
    scores = scores / temperature
    srt, idx = torch.sort(scores, descending=False, dim=-1)
    cum = srt.softmax(-1).cumsum(-1)
    remove = cum <= (1 - top_p)
    remove[..., -1:] = False
    return scores.masked_fill(remove.scatter(1, idx, remove), -float("inf"))


def greedy_decode(sd, cfg, memory, mem_mask, max_new_tokens=256, emulate_bf16=False, repetition_penalty=1.0, min_length=1,
                  sample=None):
    """Greedy restatement of vid2seq.py:150-162 with num_beams=1 (HF-4.28 generate semantics):
    start id 0, argmax, per-sequence stop at eos=1 then emit pad 0, stop when all done.  Uncached (O(S^2)).
    sample=(top_p, temperature, generator): nucleus sampling (do_sample=True) instead of the argmax."""
    ar = Arith(emulate_bf16)
    B = memory.shape[0]
    ids = torch.zeros(B, 1, dtype=torch.long, device=memory.device)
    done = torch.zeros(B, dtype=torch.bool, device=memory.device)
    d = cfg["d_model"]
    for _ in range(max_new_tokens):
        mask = torch.ones_like(ids, dtype=torch.bool)
        seq = t5_decoder(sd, cfg, ids, mask, memory, mem_mask, ar)[:, -1:] * (d ** -0.5)
        scores = process_scores(ar.linear(seq, sd["t5_model.shared.weight"])[:, 0].float(), ids, repetition_penalty, min_length)
        if sample is None:
            nxt = scores.argmax(-1)
        else:
            probs = top_p_filter(scores, sample[0], sample[1]).softmax(-1)
            nxt = torch.multinomial(probs, 1, generator=sample[2]).squeeze(1)
        nxt = torch.where(done, torch.zeros_like(nxt), nxt)
        ids = torch.cat([ids, nxt[:, None]], 1)
        done = done | (nxt == 1)
        if bool(done.all()):
            break
    return ids


def greedy_decode_cached(sd, cfg, memory, mem_mask, max_new_tokens=256, emulate_bf16=False, stop_when_done=True):
    """The same greedy loop with the reference's incremental-decoding contract (modeling_t5.py:484-525,551-556,
    1740-1766): one new token per step, per-layer cache (self-K, self-V grown by concatenation; cross-K, cross-V computed
    at step 0 and reused), relative bias of the last query row.  Same tokens as greedy_decode (tests/test_oracle_cpu.py);
    this is the form whose cost matches what `generate(use_cache=True)` does, hence bench.py's CPU decode baseline."""
    ar = Arith(emulate_bf16)
    H, dkv, d, nl = cfg["num_heads"], cfg["d_kv"], cfg["d_model"], cfg["num_layers"]
    B, E = memory.shape[0], memory.shape[1]
    dev = memory.device
    pfx = "t5_model.decoder."
    table = sd[pfx + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    cross_bias = (1.0 - mem_mask[:, None, None, :].to(torch.float32)) * NEG_MIN
    split = lambda t, L: t.view(B, L, H, dkv).transpose(1, 2)
    cross = []
    for i in range(nl):
        p = f"{pfx}block.{i}.layer.1.EncDecAttention."
        cross.append((split(ar.linear(memory, sd[p + "k.weight"]), E), split(ar.linear(memory, sd[p + "v.weight"]), E)))
    self_kv = [None] * nl
    ids = torch.zeros(B, 1, dtype=torch.long, device=dev)
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    cur = ids[:, 0]
    for step in range(max_new_tokens):
        x = sd["t5_model.shared.weight"][cur][:, None, :]                                   # (B,1,d)
        klen = step + 1
        # bias row of query position `step` over keys 0..step (modeling_t5.py:551-556: full bias, last row)
        rel = torch.arange(klen, device=dev)[None, :] - step
        bias = table[relative_position_bucket(rel, bidirectional=False)].permute(2, 0, 1).unsqueeze(0)   # (1,H,1,klen)
        for i in range(nl):
            p = f"{pfx}block.{i}."
            h = t5_layer_norm(x, sd[p + "layer.0.layer_norm.weight"])
            a = p + "layer.0.SelfAttention."
            q = split(ar.linear(h, sd[a + "q.weight"]), 1)
            k_new, v_new = split(ar.linear(h, sd[a + "k.weight"]), 1), split(ar.linear(h, sd[a + "v.weight"]), 1)
            if self_kv[i] is None:
                self_kv[i] = (k_new, v_new)
            else:
                self_kv[i] = (torch.cat([self_kv[i][0], k_new], 2), torch.cat([self_kv[i][1], v_new], 2))   # :511-515
            sc = ar.matmul(q, self_kv[i][0].transpose(3, 2)) + bias
            o = ar.softmax_pv(sc, self_kv[i][1]).transpose(1, 2).reshape(B, 1, H * dkv)
            x = x + ar.linear(o, sd[a + "o.weight"])
            h = t5_layer_norm(x, sd[p + "layer.1.layer_norm.weight"])
            a = p + "layer.1.EncDecAttention."
            q = split(ar.linear(h, sd[a + "q.weight"]), 1)
            sc = ar.matmul(q, cross[i][0].transpose(3, 2)) + cross_bias
            o = ar.softmax_pv(sc, cross[i][1]).transpose(1, 2).reshape(B, 1, H * dkv)
            x = x + ar.linear(o, sd[a + "o.weight"])
            x = t5_ff(sd, p + "layer.2.", x, ar)
        seq = t5_layer_norm(x, sd[pfx + "final_layer_norm.weight"]) * (d ** -0.5)
        nxt = ar.linear(seq, sd["t5_model.shared.weight"])[:, 0].argmax(-1)
        nxt = torch.where(done, torch.zeros_like(nxt), nxt)
        ids = torch.cat([ids, nxt[:, None]], 1)
        done = done | (nxt == 1)
        cur = nxt
        if stop_when_done and bool(done.all()):
            break
    return ids


# --------------------------------------------------------------------------------------
# Beam search (vid2seq.py:150-162 with num_beams>1 -> transformers==4.28.0 GenerationMixin.beam_search +
# BeamSearchScorer; third-party code that is NOT under /root/reference and not installable here, so this restates the
# published 4.28 algorithm; pinned against stock HF generate of transformers 5.5, tests/golden/beam_hf.pt).
# --------------------------------------------------------------------------------------
class _BeamHyps:
    """One batch item's n-best list (transformers 4.28 BeamHypotheses, early_stopping=False)."""

    def __init__(self, num_beams, length_penalty):
        self.nb, self.lp, self.beams, self.worst = num_beams, length_penalty, [], 1e9

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (hyp.shape[-1] ** self.lp)      # normalised by the length BEFORE the eos token
        if len(self.beams) < self.nb or score > self.worst:
            self.beams.append((score, hyp))
            if len(self.beams) > self.nb:
                srt = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[srt[0][1]]
                self.worst = srt[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self.beams) < self.nb:
            return False
        return self.worst >= best_sum_logprobs / cur_len ** self.lp


def beam_search_decode(sd, cfg, memory, mem_mask, num_beams=4, max_new_tokens=256, length_penalty=1.0,
                       emulate_bf16=False, eos_id=1, pad_id=0, repetition_penalty=1.0, min_length=1, num_return=1):
    """HF-4.28 beam search as called by Vid2Seq.generate: decoder_start 0, log_softmax scores, top 2*num_beams per
    batch item, eos candidates ranked below num_beams are dropped, finished hypotheses scored by
    sum_logprobs / len**length_penalty, early_stopping=False heuristic, max_new_tokens stop, best hypothesis + eos.
    Uncached decoder (O(S^2)); for small cases."""
    ar = Arith(emulate_bf16)
    B, nb, d = memory.shape[0], num_beams, cfg["d_model"]
    dev = memory.device
    mem = memory.repeat_interleave(nb, 0)
    mm = mem_mask.repeat_interleave(nb, 0)
    ids = torch.zeros(B * nb, 1, dtype=torch.long, device=dev)
    beam_scores = torch.zeros(B, nb)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_BeamHyps(nb, length_penalty) for _ in range(B)]
    done = [False] * B
    max_length = 1 + max_new_tokens
    while True:
        mask = torch.ones_like(ids, dtype=torch.bool)
        seq = t5_decoder(sd, cfg, ids, mask, mem, mm, ar)[:, -1:] * (d ** -0.5)
        logits = ar.linear(seq, sd["t5_model.shared.weight"])[:, 0].float()
        scores = process_scores(torch.log_softmax(logits, dim=-1), ids, repetition_penalty, min_length, eos_id).cpu() \
            + beam_scores[:, None]
        V = scores.shape[-1]
        top_s, top_i = torch.topk(scores.view(B, nb * V), 2 * nb, dim=1, largest=True, sorted=True)
        top_b, top_t = top_i // V, top_i % V
        cur_len = ids.shape[-1]
        n_scores = torch.zeros(B, nb)
        n_tokens = torch.zeros(B, nb, dtype=torch.long)
        n_index = torch.zeros(B, nb, dtype=torch.long)
        for b in range(B):
            if done[b]:
                n_tokens[b] = pad_id           # scores 0, indices 0 (finished batch items are padded)
                continue
            k = 0
            for rank in range(2 * nb):
                tok, sc, bi = int(top_t[b, rank]), float(top_s[b, rank]), b * nb + int(top_b[b, rank])
                if tok == eos_id:
                    if rank >= nb:
                        continue
                    hyps[b].add(ids[bi].clone().cpu(), sc)
                else:
                    n_scores[b, k], n_tokens[b, k], n_index[b, k] = sc, tok, bi
                    k += 1
                if k == nb:
                    break
            done[b] = done[b] or hyps[b].is_done(float(top_s[b].max()), cur_len)
        beam_scores = n_scores.view(-1)
        ids = torch.cat([ids[n_index.view(-1).to(dev)], n_tokens.view(-1, 1).to(dev)], 1)
        if all(done) or ids.shape[-1] >= max_length:
            break
    out = []
    for b in range(B):
        if not done[b]:
            for j in range(nb):
                hyps[b].add(ids[b * nb + j].clone().cpu(), float(beam_scores[b * nb + j]))
        ranked = sorted(hyps[b].beams, key=lambda x: x[0])
        for j in range(num_return):            # num_return_sequences: the n best hypotheses, best first
            out.append(ranked[-1 - j][1])
    sent_max = min(max(len(h) for h in out) + 1, max_length)
    dec = torch.full((B * num_return, sent_max), pad_id, dtype=torch.long)
    for b, h in enumerate(out):
        dec[b, :len(h)] = h
        if len(h) < sent_max:
            dec[b, len(h)] = eos_id
    return dec
