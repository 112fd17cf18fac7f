"""CudaOps — the op table of the Vid2Seq hot path, each op one call into libvidchap.so (include/vidchap.h).

Tensors are torch CUDA tensors used only as device-memory handles (`data_ptr()`); the arithmetic happens in the
hand-written sm_100a kernels.  All ops launch on torch's current CUDA stream.  Nothing here has a CPU path: a
missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_RELU_BWD, ACT_GELU_BWD, ACT_CE_STATS, ACT_CE_GRAD = 0, 1, 2, 3, 4, 5, 6
NO_DROP = (0, 0)   # (seed, p16): dropout spec; p16 = round(p * 65536), 0 = off


def drop_spec(p: float, seed: int):
    """(seed, p16) for a dropout site; kept values are scaled by 65536 / (65536 - p16) inside the kernels."""
    p16 = int(round(float(p) * 65536.0))
    return (int(seed) & 0xFFFFFFFF, p16) if p16 > 0 else NO_DROP


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vidchapters_b200 ops need CUDA tensors (no CPU fallback exists)")


class CudaOps:
    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("vidchapters_b200 needs a CUDA device (B200, sm_100a); none is available")
        _lib.check(self.lib.vc_device_check())
        # the device this table was built for: every launch goes to the CURRENT device's current stream, so callers
        # (Vid2Seq / Vid2SeqAdam / GraphedTrainStep entry points) make this device current first and _stream() verifies it
        self.device_index = torch.cuda.current_device()
        self.launches = 0  # kernels launched through this table (bench.py reports it as gpu_launches)

    def set_dropout_salt(self, salt):
        """salt: int32/uint32 device tensor [1] (kept alive by the caller) or None."""
        self._salt = salt
        _lib.check(self.lib.vc_set_dropout_salt(None if salt is None else C.c_void_p(salt.data_ptr())))

    def _stream(self):
        if torch.cuda.current_device() != self.device_index:
            raise RuntimeError(f"vidchapters_b200 ops built for cuda:{self.device_index} called while cuda:"
                               f"{torch.cuda.current_device()} is current (wrap the call in torch.cuda.device(...))")
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------ GEMM
    def gemm(self, A, B, out, *, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE, pre_out=None,
             aux=None, alpha=1.0, alpha_dev=None, splits=1, atomic=False, tile_n=0, drop=NO_DROP, ce=None):
        """out[M,N] = epi(alpha * op(A) @ op(B)^T).  A: [M,K] (or [K,M] if a_mn); B: [N,K] (or [K,N] if b_mn).
        ce: dict(labels, n_valid, smoothing, stats, zy | lse) for the fused LM-head + cross-entropy epilogues (act 5 / 6;
        act 5 stores nothing: out=None)."""
        _chk_cuda(A, B, out, bias, residual, pre_out, aux)
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
        assert A.dim() == 2 and B.dim() == 2 and (out is None or out.dim() == 2)
        assert A.stride(1) == 1 and B.stride(1) == 1 and (out is None or out.stride(1) == 1)
        M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
        N, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
        assert K == Kb, (A.shape, B.shape, a_mn, b_mn)
        a = _lib.GemmArgs()
        a.A, a.B = A.data_ptr(), B.data_ptr()
        a.lda, a.ldb = A.stride(0), B.stride(0)
        a.M, a.N, a.K = M, N, K
        a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
        if ce is not None:
            assert act in (ACT_CE_STATS, ACT_CE_GRAD)
            _chk_cuda(ce["labels"], ce["n_valid"], ce.get("stats"), ce.get("zy"), ce.get("lse"))
            assert ce["labels"].dtype == torch.int64 and ce["labels"].numel() == M
            a.ce_labels, a.ce_nvalid, a.ce_smoothing = ce["labels"].data_ptr(), ce["n_valid"].data_ptr(), float(ce["smoothing"])
            if act == ACT_CE_STATS:
                st = ce["stats"]
                assert tile_n in (128, 256) and st.is_contiguous() and st.dtype == torch.float32
                assert st.shape == (M, 2 * ((N + tile_n - 1) // tile_n), 3) and ce["zy"].numel() == M
                a.ce_stats, a.ce_zy = st.data_ptr(), ce["zy"].data_ptr()
            else:
                a.ce_lse = ce["lse"].data_ptr()
        if out is None:
            assert act == ACT_CE_STATS
            a.out, a.ldo, a.out_fp32 = None, (N + 7) // 8 * 8, 0
        else:
            assert out.shape[0] == M and out.shape[1] == N, (out.shape, M, N)
            a.out, a.ldo = out.data_ptr(), out.stride(0)
            a.out_fp32 = int(out.dtype == torch.float32)
            assert out.dtype in (torch.float32, torch.bfloat16)
        a.atomic = int(atomic)
        a.bias = None if bias is None else bias.data_ptr()
        if residual is not None:
            assert residual.dtype == torch.float32 and residual.stride(1) == 1
            a.residual, a.ldr = residual.data_ptr(), residual.stride(0)
        a.act = act
        if pre_out is not None:
            assert pre_out.dtype == torch.bfloat16 and pre_out.stride(0) == out.stride(0)
            a.pre_out = pre_out.data_ptr()
        if aux is not None:
            assert aux.dtype == torch.bfloat16 and aux.stride(1) == 1
            a.aux, a.ld_aux = aux.data_ptr(), aux.stride(0)
        a.alpha = float(alpha)
        a.alpha_dev = None if alpha_dev is None else alpha_dev.data_ptr()
        a.splits = int(splits)
        a.tile_n = int(tile_n)
        a.drop_seed, a.drop_p16 = drop
        _lib.check(self.lib.vc_gemm_bf16(C.byref(a), self._stream()))
        self.launches += 1
        return out

    # ------------------------------------------------------------------ attention
    def _attn_args(self, q, k, v, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel, kmask, causal, scale,
                   drop=NO_DROP, q_offset=0, q_offset_dev=None, kv_batch_rows=0, bias_zero=0, bias_len=0, q_like_k=False,
                   kv_batch_div=0):
        _chk_cuda(q, k, v, out, lse2, bias_rel, kmask)
        for t in (q, k, v, out):
            assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
        kvr = kv_batch_rows or Lk
        Bkv = B // max(kv_batch_div, 1)
        assert q.shape[0] == B * Lq and k.shape[0] == Bkv * kvr and v.shape[0] == Bkv * kvr and out.shape[0] == B * Lq
        if kmask is not None:
            assert kmask.dtype == torch.uint8 and kmask.shape == (B, Lk) and kmask.is_contiguous()
        if bias_rel is not None:
            assert bias_rel.dtype == torch.float32 and bias_rel.is_contiguous()
            assert bias_rel.shape == (H, bias_len or (Lq + Lk - 1))
        a = _lib.AttnArgs()
        a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
        a.ldq, a.ldk, a.ldv = q.stride(0), k.stride(0), v.stride(0)
        a.q_col, a.k_col, a.v_col = q_col, k_col, v_col
        a.B, a.H, a.Lq, a.Lk, a.head_dim = B, H, Lq, Lk, 64
        a.out, a.ldo = out.data_ptr(), out.stride(0)
        a.lse2 = None if lse2 is None else lse2.data_ptr()
        a.bias_rel = None if bias_rel is None else bias_rel.data_ptr()
        a.kmask = None if kmask is None else kmask.data_ptr()
        a.causal = int(causal)
        a.scale = float(scale)
        a.drop_seed, a.drop_p16 = drop
        a.q_offset, a.kv_batch_rows, a.bias_zero, a.bias_len = q_offset, kv_batch_rows, bias_zero, bias_len
        a.q_offset_dev = None if q_offset_dev is None else q_offset_dev.data_ptr()
        assert not q_like_k or (Lq == Lk and kmask is not None)
        a.q_like_k = int(bool(q_like_k))
        a.kv_batch_div = int(kv_batch_div)
        return a

    def attn_fwd(self, q, k, v, *, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel=None, kmask=None,
                 causal=False, scale=1.0, drop=NO_DROP, q_offset=0, q_offset_dev=None, kv_batch_rows=0, bias_zero=0,
                 bias_len=0, q_like_k=False, kv_batch_div=0):
        a = self._attn_args(q, k, v, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel, kmask, causal, scale, drop,
                            q_offset, q_offset_dev, kv_batch_rows, bias_zero, bias_len, q_like_k, kv_batch_div)
        _lib.check(self.lib.vc_attn_fwd(C.byref(a), self._stream()))
        self.launches += 1

    def attn_bwd(self, q, k, v, *, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel=None, kmask=None,
                 causal=False, scale=1.0, dout, do_col=0, delta, dq_acc, dk, dk_col, dv, dv_col, dbias_rel=None,
                 bucket_lut=None, drop=NO_DROP, q_like_k=False):
        """dq_acc is cleared then accumulated (fp32 atomics); dk/dv are fully written; dbias_rel accumulates."""
        _chk_cuda(dout, delta, dq_acc, dk, dv, dbias_rel, bucket_lut)
        b = _lib.AttnBwdArgs()
        b.fwd = self._attn_args(q, k, v, q_col, k_col, v_col, B, H, Lq, Lk, out, lse2, bias_rel, kmask, causal, scale,
                                drop, q_like_k=q_like_k)
        assert dout.dtype == torch.bfloat16 and dq_acc.dtype == torch.float32 and delta.dtype == torch.float32
        assert dk.dtype == torch.bfloat16 and dv.dtype == torch.bfloat16
        b.dout, b.ld_do, b.do_col = dout.data_ptr(), dout.stride(0), do_col
        b.delta = delta.data_ptr()
        b.dq_acc, b.ld_dq = dq_acc.data_ptr(), dq_acc.stride(0)
        b.dk, b.ld_dk, b.dk_col = dk.data_ptr(), dk.stride(0), dk_col
        b.dv, b.ld_dv, b.dv_col = dv.data_ptr(), dv.stride(0), dv_col
        b.dbias_rel = None if dbias_rel is None else dbias_rel.data_ptr()
        if bucket_lut is not None:
            assert bucket_lut.dtype == torch.int32
        b.bucket_lut = None if bucket_lut is None else bucket_lut.data_ptr()
        _lib.check(self.lib.vc_attn_bwd(C.byref(b), self._stream()))
        self.launches += 2

    # ------------------------------------------------------------------ norms
    def norm_fwd(self, kind, x, w, bias, *, out_bf16=None, out_f32=None, rstd=None, mean=None, eps, out_scale=1.0,
                 rows_per_batch=0, out_batch_stride=0, out_row_offset=0, drop=NO_DROP):
        _chk_cuda(x, w, bias, out_bf16, out_f32, rstd, mean)
        M, D = x.shape
        assert x.dtype == torch.float32 and x.is_contiguous()
        _lib.check(self.lib.vc_norm_fwd(kind, _ptr(x), _ptr(w), _ptr(bias), _ptr(out_bf16), _ptr(out_f32), _ptr(rstd),
                                        _ptr(mean), M, D, eps, out_scale, rows_per_batch, out_batch_stride,
                                        out_row_offset, drop[0], drop[1], self._stream()))
        self.launches += 1

    def norm_bwd(self, kind, g, x, w, rstd, mean, *, dx, dx_bf16=None, accumulate_dx, dw, db=None, scale=1.0,
                 rows_per_batch=0, g_batch_stride=0, g_row_offset=0, g_drop=NO_DROP, dxb_drop=NO_DROP):
        _chk_cuda(g, x, w, rstd, mean, dx, dx_bf16, dw, db)
        M, D = x.shape
        assert g.dtype in (torch.float32, torch.bfloat16) and g.is_contiguous()
        assert dx.dtype == torch.float32 and x.is_contiguous() and dx.is_contiguous()
        _lib.check(self.lib.vc_norm_bwd(kind, _ptr(g), int(g.dtype == torch.bfloat16), _ptr(x), _ptr(w), _ptr(rstd), _ptr(mean), _ptr(dx), _ptr(dx_bf16),
                                        int(accumulate_dx), _ptr(dw), _ptr(db), M, D, scale, rows_per_batch,
                                        g_batch_stride, g_row_offset, g_drop[0], g_drop[1], dxb_drop[0], dxb_drop[1],
                                        self._stream()))
        self.launches += 1

    # ------------------------------------------------------------------ small ops
    def embed_fwd(self, ids, table, out, drop=NO_DROP):
        _chk_cuda(ids, table, out)
        assert ids.dtype == torch.int64 and ids.is_contiguous() and out.is_contiguous()
        _lib.check(self.lib.vc_embed_fwd(_ptr(ids), _ptr(table), _ptr(out), ids.numel(), table.shape[1], table.shape[0],
                                         drop[0], drop[1], self._stream()))
        self.launches += 1

    def embed_bwd(self, ids, dout, dtable, drop=NO_DROP):
        _chk_cuda(ids, dout, dtable)
        _lib.check(self.lib.vc_embed_bwd(_ptr(ids), _ptr(dout), _ptr(dtable), ids.numel(), dtable.shape[1],
                                         dtable.shape[0], drop[0], drop[1], self._stream()))
        self.launches += 1

    def prepare_targets(self, out_ids, dec_in, labels, n_valid, pad_id=0):
        _chk_cuda(out_ids, dec_in, labels, n_valid)
        B, S = out_ids.shape
        assert out_ids.dtype == torch.int64 and out_ids.is_contiguous()
        _lib.check(self.lib.vc_prepare_targets(_ptr(out_ids), _ptr(dec_in), _ptr(labels), _ptr(n_valid), B, S, pad_id,
                                               self._stream()))
        self.launches += 1

    def bias_expand(self, table, lut, out):
        _chk_cuda(table, lut, out)
        H, R = out.shape
        _lib.check(self.lib.vc_bias_expand(_ptr(table), _ptr(lut), _ptr(out), H, R, self._stream()))
        self.launches += 1

    def bias_fold(self, drel, lut, dtable):
        _chk_cuda(drel, lut, dtable)
        H, R = drel.shape
        _lib.check(self.lib.vc_bias_fold(_ptr(drel), _ptr(lut), _ptr(dtable), H, R, self._stream()))
        self.launches += 1

    def add_pos(self, x, pos, out, P, drop=NO_DROP):
        _chk_cuda(x, pos, out)
        B, T, Cc = x.shape
        assert x.dtype == torch.float32 and x.is_contiguous()
        _lib.check(self.lib.vc_add_pos(_ptr(x), _ptr(pos), _ptr(out), B, T, Cc, P, drop[0], drop[1], self._stream()))
        self.launches += 1

    def add_pos_bwd(self, dx, dpos, B, T, Cc, P, drop=NO_DROP):
        _chk_cuda(dx, dpos)
        _lib.check(self.lib.vc_add_pos_bwd(_ptr(dx), _ptr(dpos), B, T, Cc, P, drop[0], drop[1], self._stream()))
        self.launches += 1

    def cross_entropy(self, logits, labels, n_valid, smoothing, loss_out, dlogits):
        _chk_cuda(logits, labels, n_valid, loss_out, dlogits)
        n, V = logits.shape
        assert logits.dtype == torch.float32 and logits.stride(1) == 1
        _lib.check(self.lib.vc_cross_entropy(_ptr(logits), logits.stride(0), _ptr(labels), _ptr(n_valid), smoothing,
                                             _ptr(loss_out), _ptr(dlogits), 0 if dlogits is None else dlogits.stride(0),
                                             n, V, self._stream()))
        self.launches += 1

    def ce_combine(self, stats, zy, labels, n_valid, smoothing, V, lse_out, loss_out):
        """Partials of the act-5 LM-head GEMM -> lse [M] and the label-smoothed mean loss (modeling_t5.py:1721)."""
        _chk_cuda(stats, zy, labels, n_valid, lse_out, loss_out)
        M, n_slots, _ = stats.shape
        _lib.check(self.lib.vc_ce_combine(_ptr(stats), n_slots, _ptr(zy), _ptr(labels), _ptr(n_valid), smoothing, V,
                                          _ptr(lse_out), _ptr(loss_out), M, self._stream()))
        self.launches += 1

    def colsum_bf16(self, x, out):
        _chk_cuda(x, out)
        M, N = x.shape
        _lib.check(self.lib.vc_colsum_bf16(_ptr(x), x.stride(0), _ptr(out), M, N, self._stream()))
        self.launches += 1

    def cast_f32_bf16(self, src, dst, scale=1.0):
        _chk_cuda(src, dst)
        M, N = src.shape
        _lib.check(self.lib.vc_cast_f32_bf16(_ptr(src), src.stride(0), _ptr(dst), dst.stride(0), M, N, scale,
                                             self._stream()))
        self.launches += 1

    def copy_rows_bf16(self, src, dst, B, T, Cc, E, row_off):
        _chk_cuda(src, dst)
        _lib.check(self.lib.vc_copy_rows_bf16(_ptr(src), _ptr(dst), B, T, Cc, E, row_off, self._stream()))
        self.launches += 1

    # ------------------------------------------------------------------ incremental decoding
    def kv_append(self, src, cache, pos_dev):
        """cache[b, *pos_dev, :] = src[b, :]; src [B, C] bf16 (any row stride), cache [B, cap, C]."""
        _chk_cuda(src, cache, pos_dev)
        B, cap, Cc = cache.shape
        assert src.shape == (B, Cc) and src.stride(1) == 1 and cache.is_contiguous() and pos_dev.dtype == torch.int32
        _lib.check(self.lib.vc_kv_append(_ptr(src), src.stride(0), _ptr(cache), B, cap, Cc, _ptr(pos_dev), self._stream()))
        self.launches += 1

    def decode_linear(self, A, W, out, *, norm_w=None, eps=1e-6, out_scale=1.0, residual=None, relu=False):
        """out[M,N] = epi(A[M,K] @ W[N,K]^T) for the M = batch rows of a decode step; A bf16, or the fp32 residual stream
        with the T5 RMS norm fused (norm_w); relu / + residual (fp32 out, may alias out) epilogues."""
        _chk_cuda(A, W, out, norm_w, residual)
        M, K = A.shape
        N = W.shape[0]
        assert W.dtype == torch.bfloat16 and W.shape[1] == K and W.stride(1) == 1 and A.stride(1) == 1 and out.stride(1) == 1
        assert A.dtype in (torch.float32, torch.bfloat16) and out.dtype in (torch.float32, torch.bfloat16)
        assert out.shape == (M, N) and (residual is None or (residual.dtype == torch.float32 and residual.stride(1) == 1))
        _lib.check(self.lib.vc_decode_linear(_ptr(A), A.stride(0), int(A.dtype == torch.float32), _ptr(norm_w), eps, out_scale,
                                             _ptr(W), W.stride(0), _ptr(out), out.stride(0), int(out.dtype == torch.float32),
                                             _ptr(residual), 0 if residual is None else residual.stride(0), int(relu), M, N,
                                             K, self._stream()))
        self.launches += 1

    def greedy_next(self, logits, done, ids_out, seq, pos_dev, eos_id=1, pad_id=0):
        _chk_cuda(logits, done, ids_out, seq, pos_dev)
        B, V = logits.shape
        assert logits.dtype == torch.float32 and logits.stride(1) == 1 and done.dtype == torch.uint8
        assert ids_out.dtype == torch.int64 and seq.dtype == torch.int64 and seq.is_contiguous()
        _lib.check(self.lib.vc_greedy_next(_ptr(logits), logits.stride(0), V, _ptr(done), _ptr(ids_out), _ptr(seq),
                                           seq.shape[1], _ptr(pos_dev), eos_id, pad_id, B, self._stream()))
        self.launches += 1

    def beam_topk(self, logits, beam_scores, num_beams, out_scores, out_tokens, out_beams):
        """Top 2*num_beams of log_softmax(logits) + beam_scores per batch item (HF-4.28 beam_search candidate step)."""
        _chk_cuda(logits, beam_scores, out_scores, out_tokens, out_beams)
        Bn, V = logits.shape
        B = Bn // num_beams
        assert logits.dtype == torch.float32 and logits.stride(1) == 1 and beam_scores.dtype == torch.float32
        assert out_scores.shape == (B, 2 * num_beams) and out_scores.is_contiguous() and out_scores.dtype == torch.float32
        assert out_tokens.dtype == torch.int32 and out_beams.dtype == torch.int32
        assert out_tokens.is_contiguous() and out_beams.is_contiguous() and beam_scores.is_contiguous()
        _lib.check(self.lib.vc_beam_topk(_ptr(logits), logits.stride(0), V, _ptr(beam_scores), num_beams, B, _ptr(out_scores),
                                         _ptr(out_tokens), _ptr(out_beams), self._stream()))
        self.launches += 1

    def kv_reorder(self, src, dst, beam_idx, n):
        """dst[b, :n] = src[beam_idx[b], :n] for bf16 caches [Bn, cap, C] (HF _reorder_cache)."""
        _chk_cuda(src, dst, beam_idx)
        Bn, cap, C = src.shape
        assert src.dtype == torch.bfloat16 and dst.shape == src.shape and src.is_contiguous() and dst.is_contiguous()
        assert beam_idx.dtype == torch.int32 and beam_idx.numel() == Bn
        _lib.check(self.lib.vc_kv_reorder(_ptr(src), _ptr(dst), _ptr(beam_idx), Bn, cap, C, int(n), self._stream()))
        self.launches += 1

    def step_advance(self, pos_dev):
        _chk_cuda(pos_dev)
        _lib.check(self.lib.vc_step_advance(_ptr(pos_dev), self._stream()))
        self.launches += 1

    # ------------------------------------------------------------------ optimiser tail
    def sumsq(self, g, out_accum):
        _chk_cuda(g, out_accum)
        _lib.check(self.lib.vc_sumsq(_ptr(g), g.numel(), _ptr(out_accum), self._stream()))
        self.launches += 1

    def adam_step(self, p, g, m, v, p_bf16, *, lr, beta1, beta2, eps, step, norm_sq=None, clip_max_norm=0.0,
                  grad_scale=1.0):
        _chk_cuda(p, g, m, v, p_bf16, norm_sq)
        _lib.check(self.lib.vc_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), p.numel(), lr, beta1, beta2, eps,
                                         step, _ptr(norm_sq), clip_max_norm, grad_scale, self._stream()))
        self.launches += 1

    def renorm_time_tokens(self, w, w_bf16, num_bins, scratch2):
        _chk_cuda(w, w_bf16, scratch2)
        V, d = w.shape
        _lib.check(self.lib.vc_renorm_time_tokens(_ptr(w), _ptr(w_bf16), V, d, num_bins, _ptr(scratch2), self._stream()))
        self.launches += 2

    def cast_flat_bf16(self, src, dst):
        _chk_cuda(src, dst)
        _lib.check(self.lib.vc_cast_flat_bf16(_ptr(src), _ptr(dst), src.numel(), self._stream()))
        self.launches += 1
