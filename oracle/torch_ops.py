"""TEST INFRASTRUCTURE — per-op oracle: the CudaOps table (vidchapters_b200/ops.py) restated with plain torch ops.

Used only by tests/ (and __graft_entry__.smoke()):
  * on the GPU, as the "plain PyTorch fp32 reference of the same op" each CUDA kernel is compared against;
  * on CPU, injected into the host orchestration (vidchapters_b200.engine) so that the whole forward/backward WIRING
    can be checked against oracle/vid2seq_oracle.py (and through it the real reference) without a GPU.
The product never selects it: vidchapters_b200.Vid2Seq builds CudaOps itself and raises without a B200.

Rounding points mirror the kernels: GEMM operands and stored bf16 tensors are bf16, accumulation fp32.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_GELU, ACT_RELU_BWD, ACT_GELU_BWD, ACT_CE_STATS, ACT_CE_GRAD = 0, 1, 2, 3, 4, 5, 6
_MASKED = -3.0e38


NO_DROP = (0, 0)
_M32 = 0xFFFFFFFF
_SALT = None   # optional device/host tensor XORed into every seed (mirrors vc_set_dropout_salt)


def _lane_mult(i: torch.Tensor) -> torch.Tensor:
    """A[i] of vidchapters_b200/csrc/ptx.cuh::drop_lane_mult (fixed odd multiplier of in-block column i)."""
    x = ((i + 1) * 0x9E3779B1) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x85EBCA6B) & _M32
    x = x ^ (x >> 13)
    return x | 1


def drop_mask(spec, idx: torch.Tensor, scaled: bool = True) -> torch.Tensor:
    """Float multiplier (0 or 65536/(65536-p16); 0/1 with scaled=False) for every element of a tensor whose flat int64
    indices are `idx` (shape [..., C]): the counter-based mask of vidchapters_b200/csrc/ptx.cuh (row = flat // C, column
    = flat % C; one strong odd hash H per 32-column block, element value = top 16 bits of H * A[column % 32])."""
    seed, p16 = spec
    if p16 == 0:
        return torch.ones(idx.shape, dtype=torch.float32, device=idx.device)
    if _SALT is not None:
        seed = (seed ^ (int(_SALT.item()) & _M32)) & _M32
    C = idx.shape[-1]
    row, col = idx // C, idx % C
    key = (seed + (row & _M32) * 0x9E3779B1 + ((row >> 32) & _M32) * 0x7F4A7C15) & _M32
    x = (key + (col >> 5) * 0x85EBCA77) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x2C1B3C6D) & _M32
    x = x ^ (x >> 12)
    x = (x * 0x297A2D39) & _M32
    x = x ^ (x >> 15)
    h = x | 1
    r = ((h * _lane_mult(col & 31)) & _M32) >> 16
    return (r >= p16).float() * ((65536.0 / (65536.0 - p16)) if scaled else 1.0)


def _idx(shape, device):
    n = 1
    for s_ in shape:
        n *= s_
    return torch.arange(n, dtype=torch.int64, device=device).view(shape)


def _gelu_grad(x):
    cdf = 0.5 * (1.0 + torch.erf(x * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * x * x)
    return cdf + x * pdf


class TorchOps:
    name = "torch-oracle"

    def __init__(self, flash_rounding: bool = True):
        self.launches = 0
        # flash_rounding: round the UN-normalised probabilities 2^(s2 - ceil(rowmax2)) to bf16 before P.V and divide by
        # the row sum afterwards — the attention kernel's rounding points (see oracle/vid2seq_oracle.py::Arith).
        self.flash = flash_rounding

    def set_dropout_salt(self, salt):
        global _SALT
        _SALT = salt

    # ------------------------------------------------------------------ GEMM
    def gemm(self, A, B, out, *, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE, pre_out=None,
             aux=None, alpha=1.0, alpha_dev=None, splits=1, atomic=False, tile_n=0, drop=NO_DROP, ce=None):
        a = A.float().t() if a_mn else A.float()
        b = B.float() if b_mn else B.float().t()
        acc = (a @ b) * alpha
        if act == ACT_CE_STATS:      # fused LM head + CE, statistics pass: per-row (max, sum exp, sum z) in slot 0
            st, lab = ce["stats"], ce["labels"].reshape(-1)
            st.zero_()
            st[:, :, 0] = float("-inf")
            m = acc.max(-1).values
            st[:, 0, 0], st[:, 0, 1], st[:, 0, 2] = m, torch.exp(acc - m[:, None]).sum(-1), acc.sum(-1)
            ce["zy"].copy_(acc.gather(1, lab.clamp(min=0)[:, None]).squeeze(1))
            return None
        if act == ACT_CE_GRAD:
            lab = ce["labels"].reshape(-1)
            g = torch.exp(acc - ce["lse"][:, None]) - ce["smoothing"] / acc.shape[1]
            g[torch.arange(acc.shape[0], device=g.device), lab.clamp(min=0)] -= (1 - ce["smoothing"])
            g = g * (lab != -100)[:, None] / ce["n_valid"].reshape(())
            out.copy_(g.to(out.dtype))
            return out
        if alpha_dev is not None:
            acc = acc * alpha_dev.float()
        if bias is not None:
            acc = acc + bias
        if act == ACT_GELU and pre_out is not None:
            pre_out.copy_(acc.to(pre_out.dtype))
        if act == ACT_RELU:
            acc = torch.relu(acc)
        elif act == ACT_GELU:
            acc = F.gelu(acc)
        elif act == ACT_RELU_BWD:
            acc = acc * (aux.float() > 0)
        elif act == ACT_GELU_BWD:
            acc = acc * _gelu_grad(aux.float())
        if drop[1]:
            acc = acc * drop_mask(drop, _idx(acc.shape, acc.device))
        if residual is not None:
            acc = acc + residual
        if atomic:
            out.add_(acc.to(out.dtype))
        else:
            out.copy_(acc.to(out.dtype))
        return out

    # ------------------------------------------------------------------ attention
    @staticmethod
    def _heads(buf, col, B, L, H):
        return buf[:, col:col + H * 64].float().reshape(B, L, H, 64).permute(0, 2, 1, 3)  # B,H,L,64

    def _scores(self, q, k, q_col, k_col, B, H, Lq, Lk, bias_rel, kmask, causal, scale, qoff=0, kv_rows=0, bias_zero=None):
        qh = self._heads(q, q_col, B, Lq, H)
        kvr = kv_rows or Lk
        kh = k[:, k_col:k_col + H * 64].float().reshape(B, kvr, H, 64)[:, :Lk].permute(0, 2, 1, 3)
        s = (qh @ kh.transpose(-1, -2)) * scale
        qpos = torch.arange(Lq, device=s.device)[:, None] + qoff
        if bias_rel is not None:
            bz = (Lq - 1) if bias_zero is None else bias_zero
            idx = (torch.arange(Lk, device=s.device)[None, :] - qpos) + bz
            s = s + bias_rel[:, idx][None]
        masked = torch.zeros(B, 1, Lq, Lk, dtype=torch.bool, device=s.device)
        if kmask is not None:
            masked = masked | (kmask[:, None, None, :] == 0)
        if causal:
            masked = masked | (torch.arange(Lk, device=s.device)[None, :] > qpos)[None, None]
        return torch.where(masked, torch.full_like(s, _MASKED), s)

    def attn_fwd(self, q, k, v, *, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel=None, kmask=None,
                 causal=False, scale=1.0, drop=NO_DROP, q_offset=0, q_offset_dev=None, kv_batch_rows=0, bias_zero=0,
                 bias_len=0, q_like_k=False, kv_batch_div=0):
        if kv_batch_div > 1:      # beams of one video share its K/V: expand for this checker
            kvr_ = kv_batch_rows or Lk
            k = k.view(B // kv_batch_div, kvr_, -1).repeat_interleave(kv_batch_div, 0).reshape(B * kvr_, -1)
            v = v.view(B // kv_batch_div, kvr_, -1).repeat_interleave(kv_batch_div, 0).reshape(B * kvr_, -1)
        # q_like_k (skip padding-only query tiles) is a pure work-skipping hint: this checker computes every row
        qoff = q_offset + (int(q_offset_dev.item()) if q_offset_dev is not None else 0)
        s = self._scores(q, k, q_col, k_col, B, H, Lq, Lk, bias_rel, kmask, causal, scale, qoff, kv_batch_rows,
                         bias_zero if bias_len else None)
        kvr = kv_batch_rows or Lk
        vh = v[:, v_col:v_col + H * 64].float().reshape(B, kvr, H, 64)[:, :Lk].permute(0, 2, 1, 3)
        # kernel rounding points: the kept probabilities are rounded to bf16 UNSCALED; 1/(1-p) multiplies the output
        dm = drop_mask(drop, _idx(s.shape, s.device), scaled=False) if drop[1] else 1.0
        sc = 65536.0 / (65536.0 - drop[1]) if drop[1] else 1.0
        if self.flash:
            s2 = s * math.log2(math.e)
            e = torch.exp2(s2 - torch.ceil(s2.max(-1, keepdim=True).values))
            o = ((e * dm).to(torch.bfloat16).float() @ vh) / e.sum(-1, keepdim=True)
            o = o * sc if drop[1] else o
        else:
            o = ((torch.softmax(s, dim=-1) * dm).to(torch.bfloat16).float() @ vh) * sc
        out[:, :H * 64].copy_(o.permute(0, 2, 1, 3).reshape(B * Lq, H * 64).to(out.dtype))
        if lse2 is not None:
            lse2.copy_(torch.logsumexp(s, dim=-1) * math.log2(math.e))

    def attn_bwd(self, q, k, v, *, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel=None, kmask=None,
                 causal=False, scale=1.0, dout, do_col=0, delta, dq_acc, dk, dk_col, dv, dv_col, dbias_rel=None,
                 bucket_lut=None, drop=NO_DROP, q_like_k=False):
        s = self._scores(q, k, q_col, k_col, B, H, Lq, Lk, bias_rel, kmask, causal, scale)
        p = torch.softmax(s, dim=-1)
        dm = drop_mask(drop, _idx(p.shape, p.device)) if drop[1] else None
        qh, kh, vh = (self._heads(q, q_col, B, Lq, H), self._heads(k, k_col, B, Lk, H), self._heads(v, v_col, B, Lk, H))
        do = self._heads(dout, do_col, B, Lq, H)
        oh = self._heads(out, 0, B, Lq, H)
        dlt = (do * oh).sum(-1, keepdim=True)
        delta.copy_(dlt.squeeze(-1))
        dp = do @ vh.transpose(-1, -2)
        if dm is not None:
            dp = dp * dm
        ds = p * (dp - dlt)
        sc = 65536.0 / (65536.0 - drop[1]) if drop[1] else 1.0
        pb = (p if dm is None else p * (dm > 0)).to(torch.bfloat16).float()   # masked, unscaled (as the kernel)
        dsb = (ds * scale).to(torch.bfloat16).float()
        dvh = (pb.transpose(-1, -2) @ do) * sc
        dkh = dsb.transpose(-1, -2) @ qh
        dqh = dsb @ kh
        dq_acc[:, :H * 64].copy_(dqh.permute(0, 2, 1, 3).reshape(B * Lq, H * 64))
        dk[:, dk_col:dk_col + H * 64].copy_(dkh.permute(0, 2, 1, 3).reshape(B * Lk, H * 64).to(dk.dtype))
        dv[:, dv_col:dv_col + H * 64].copy_(dvh.permute(0, 2, 1, 3).reshape(B * Lk, H * 64).to(dv.dtype))
        if dbias_rel is not None:
            idx = (torch.arange(Lk, device=s.device)[None, :] - torch.arange(Lq, device=s.device)[:, None]) + Lq - 1
            g = ds.sum(0)  # H,Lq,Lk
            dbias_rel.index_put_((torch.arange(H, device=s.device)[:, None, None].expand(H, Lq, Lk),
                                  idx[None].expand(H, Lq, Lk)), g, accumulate=True)

    # ------------------------------------------------------------------ norms
    @staticmethod
    def _rows(M, rows_per_batch, batch_stride, row_offset, device):
        r = torch.arange(M, device=device)
        if rows_per_batch > 0:
            return (r // rows_per_batch) * batch_stride + row_offset + r % rows_per_batch
        return r

    def norm_fwd(self, kind, x, w, bias, *, out_bf16=None, out_f32=None, rstd=None, mean=None, eps, out_scale=1.0,
                 rows_per_batch=0, out_batch_stride=0, out_row_offset=0, drop=NO_DROP):
        M, D = x.shape
        if kind == 1:
            mu = x.mean(-1, keepdim=True)
            var = ((x - mu) ** 2).mean(-1, keepdim=True)
        else:
            mu = torch.zeros(M, 1, device=x.device)
            var = (x ** 2).mean(-1, keepdim=True)
        rs = torch.rsqrt(var + eps)
        y = (x - mu) * rs * w
        if kind == 1:
            y = y + bias
        y = y * out_scale
        if drop[1]:
            y = y * drop_mask(drop, _idx(y.shape, y.device))
        rows = self._rows(M, rows_per_batch, out_batch_stride, out_row_offset, x.device)
        if out_bf16 is not None:
            out_bf16.view(-1, D)[rows] = y.to(out_bf16.dtype)
        if out_f32 is not None:
            out_f32.view(-1, D)[rows] = y
        if rstd is not None:
            rstd.copy_(rs.squeeze(-1))
        if mean is not None and kind == 1:
            mean.copy_(mu.squeeze(-1))

    def norm_bwd(self, kind, g, x, w, rstd, mean, *, dx, dx_bf16=None, accumulate_dx, dw, db=None, scale=1.0,
                 rows_per_batch=0, g_batch_stride=0, g_row_offset=0, g_drop=NO_DROP, dxb_drop=NO_DROP):
        M, D = x.shape
        rows = self._rows(M, rows_per_batch, g_batch_stride, g_row_offset, x.device)
        gg = g.view(-1, D)[rows].float() * scale
        if g_drop[1]:
            gg = gg * drop_mask(g_drop, _idx(gg.shape, gg.device))
        mu = mean[:, None] if kind == 1 else 0.0
        xh = (x - mu) * rstd[:, None]
        dh = gg * w
        m2 = (dh * xh).mean(-1, keepdim=True)
        m1 = dh.mean(-1, keepdim=True) if kind == 1 else 0.0
        d = rstd[:, None] * (dh - m1 - xh * m2)
        if accumulate_dx:
            dx.add_(d)
        else:
            dx.copy_(d)
        if dx_bf16 is not None:
            dd = dx * drop_mask(dxb_drop, _idx(dx.shape, dx.device)) if dxb_drop[1] else dx
            dx_bf16.copy_(dd.to(dx_bf16.dtype))
        if dw is not None:
            dw.add_((gg * xh).sum(0))
        if db is not None and kind == 1:
            db.add_(gg.sum(0))

    # ------------------------------------------------------------------ small ops
    def embed_fwd(self, ids, table, out, drop=NO_DROP):
        e = table[ids.reshape(-1)].view_as(out)
        out.copy_(e * drop_mask(drop, _idx(e.shape, e.device)) if drop[1] else e)

    def embed_bwd(self, ids, dout, dtable, drop=NO_DROP):
        g = dout.reshape(-1, dtable.shape[1])
        if drop[1]:
            g = g * drop_mask(drop, _idx(g.shape, g.device))
        dtable.index_add_(0, ids.reshape(-1), g)

    def prepare_targets(self, out_ids, dec_in, labels, n_valid, pad_id=0):
        lab = out_ids.masked_fill(out_ids == pad_id, -100)
        labels.copy_(lab)
        s = torch.zeros_like(out_ids)
        s[:, 1:] = lab[:, :-1]
        dec_in.copy_(s.masked_fill(s == -100, 0))
        n_valid.fill_(float((lab != -100).sum()))

    def bias_expand(self, table, lut, out):
        out.copy_(table[lut.long()].t())

    def bias_fold(self, drel, lut, dtable):
        dtable.index_add_(0, lut.long(), drel.t().contiguous())

    @staticmethod
    def _pos_idx(T, P, device):
        if T == P:
            return torch.arange(T, device=device)
        return torch.floor(torch.arange(T, device=device, dtype=torch.float32) * (float(P) / float(T))).long()

    def add_pos(self, x, pos, out, P, drop=NO_DROP):
        B, T, C = x.shape
        y = x + pos.view(P, C)[self._pos_idx(T, P, x.device)][None]
        out.copy_(y * drop_mask(drop, _idx(y.shape, y.device)) if drop[1] else y)

    def add_pos_bwd(self, dx, dpos, B, T, C, P, drop=NO_DROP):
        g = dx.view(B, T, C)
        if drop[1]:
            g = g * drop_mask(drop, _idx(g.shape, g.device))
        dpos.view(P, C).index_add_(0, self._pos_idx(T, P, dx.device), g.sum(0))

    def cross_entropy(self, logits, labels, n_valid, smoothing, loss_out, dlogits):
        n, V = logits.shape
        lab = labels.reshape(-1)
        valid = lab != -100
        lp = torch.log_softmax(logits.float(), dim=-1)
        nll = -lp.gather(1, lab.clamp(min=0)[:, None]).squeeze(1)
        smooth = -lp.mean(-1)
        row = (1 - smoothing) * nll + smoothing * smooth
        nv = n_valid.reshape(())
        loss_out.fill_(0.0)
        loss_out.add_((row * valid).sum() / nv)
        if dlogits is not None:
            g = lp.exp() - smoothing / V
            g[torch.arange(n, device=g.device), lab.clamp(min=0)] -= (1 - smoothing)
            g = g * valid[:, None] / nv
            dlogits.copy_(g.to(dlogits.dtype))

    def ce_combine(self, stats, zy, labels, n_valid, smoothing, V, lse_out, loss_out):
        m = stats[:, :, 0].max(-1).values
        s = (stats[:, :, 1] * torch.exp(stats[:, :, 0] - m[:, None])).sum(-1)
        t = stats[:, :, 2].sum(-1)
        lse = m + torch.log(s)
        lse_out.copy_(lse)
        lab = labels.reshape(-1)
        row = (1 - smoothing) * (lse - zy) + smoothing * (lse - t / V)
        loss_out.fill_(0.0)
        loss_out.add_((row * (lab != -100)).sum() / n_valid.reshape(()))

    def colsum_bf16(self, x, out):
        out.add_(x.float().sum(0))

    def cast_f32_bf16(self, src, dst, scale=1.0):
        dst.copy_((src * scale).to(dst.dtype))

    def copy_rows_bf16(self, src, dst, B, T, C, E, row_off):
        dst.view(B, E, C)[:, row_off:row_off + T].copy_(src.view(B, T, C))

    # ------------------------------------------------------------------ incremental decoding
    def kv_append(self, src, cache, pos_dev):
        cache[:, int(pos_dev.item())].copy_(src)

    def decode_linear(self, A, W, out, *, norm_w=None, eps=1e-6, out_scale=1.0, residual=None, relu=False):
        a = A.float()
        if norm_w is not None:
            a = a * torch.rsqrt((a * a).mean(-1, keepdim=True) + eps) * out_scale * norm_w
        acc = a.to(torch.bfloat16).float() @ W.float().t()
        if relu:
            acc = torch.relu(acc)
        if residual is not None:
            acc = acc + residual
        out.copy_(acc.to(out.dtype))

    def greedy_next(self, logits, done, ids_out, seq, pos_dev, eos_id=1, pad_id=0):
        nxt = logits.argmax(-1)
        nxt = torch.where(done.bool(), torch.full_like(nxt, pad_id), nxt)
        done.copy_((done.bool() | (nxt == eos_id)).to(torch.uint8))
        ids_out.copy_(nxt)
        pos = int(pos_dev.item())
        if pos + 1 < seq.shape[1]:
            seq[:, pos + 1] = nxt

    def beam_topk(self, logits, beam_scores, num_beams, out_scores, out_tokens, out_beams):
        Bn, V = logits.shape
        B = Bn // num_beams
        sc = torch.log_softmax(logits.float(), dim=-1) + beam_scores.view(-1, 1)
        s, i = torch.topk(sc.view(B, num_beams * V), 2 * num_beams, dim=1, largest=True, sorted=True)
        out_scores.copy_(s)
        out_tokens.copy_((i % V).to(torch.int32))
        out_beams.copy_((i // V).to(torch.int32))

    def kv_reorder(self, src, dst, beam_idx, n):
        dst[:, :n].copy_(src[beam_idx.long(), :n])

    def step_advance(self, pos_dev):
        pos_dev.add_(1)

    # ------------------------------------------------------------------ optimiser tail
    def sumsq(self, g, out_accum):
        out_accum.add_((g.double() ** 2).sum().float())

    def adam_step(self, p, g, m, v, p_bf16, *, lr, beta1, beta2, eps, step, norm_sq=None, clip_max_norm=0.0,
                  grad_scale=1.0):
        coef = grad_scale
        if norm_sq is not None and clip_max_norm > 0:
            total = torch.sqrt(norm_sq.reshape(())) * grad_scale
            coef = coef * torch.clamp(clip_max_norm / (total + 1e-6), max=1.0)
        gg = g * coef
        m.mul_(beta1).add_(gg, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
        if p_bf16 is not None:
            p_bf16.copy_(p.to(p_bf16.dtype))

    def renorm_time_tokens(self, w, w_bf16, num_bins, scratch2):
        frozen = torch.norm(w[:-num_bins], dim=1).mean(0)
        train = torch.norm(w[-num_bins:], dim=1).mean(0)
        w[-num_bins:].div_(train / frozen)
        if w_bf16 is not None:
            w_bf16[-num_bins:].copy_(w[-num_bins:].to(w_bf16.dtype))

    def cast_flat_bf16(self, src, dst):
        dst.copy_(src.to(dst.dtype))
