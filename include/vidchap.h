/* vidchap.h — C ABI of libvidchap.so: the hand-written sm_100a kernels under vidchapters_b200.Vid2Seq.
 *
 * The reference (antoyang/VidChapters) has NO FFI/plugin interface on this path: its hot path is PyTorch
 * library calls inside model/vid2seq.py, model/vit.py, model/modeling_t5.py and dvc.py:112-126.  The drop-in
 * boundary is therefore the Python class contract; this C ABI is what that class binds through
 * ctypes, one entry point per library call it replaces.  Each declaration cites the reference lines replaced.
 *
 * Conventions: all pointers are DEVICE pointers unless named host_*; `stream` is a cudaStream_t passed as void*;
 * every function returns VC_OK (0) or a negative status, never aborts; vc_last_error() returns the message of the
 * calling thread's last failure.  No function falls back to a CPU path.
 */
#ifndef VIDCHAP_H_
#define VIDCHAP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC_OK 0
#define VC_ERR_INVALID (-1)
#define VC_ERR_CUDA (-2)

int vc_version(void);               /* ABI version, bumps on any signature change */
const char* vc_last_error(void);    /* thread-local message of the last failure */
int vc_device_check(void);          /* VC_OK iff the current device is sm_100 (B200); message otherwise */
/* Device pointer to a 32-bit salt XORed into every dropout seed by every kernel launched afterwards (NULL = none):
 * a CUDA graph captured once draws fresh masks on each replay when the host bumps the salt between replays. */
int vc_set_dropout_salt(const uint32_t* dev_ptr);
/* Debug only: device buffer of >= 2048 int64; the attention kernels then record a clock64 timeline of their CTA (0,0,0)
 * (event id in the top 16 bits).  NULL (default) turns it off.  Used by tools/attn_timeline.py. */
int vc_debug_set_trace(void* dev_ptr);

/* ---- GEMM: out[M,N] = epilogue( alpha * op(A)[M,K] . op(B)[N,K]^T ), bf16 operands, fp32 accumulate (tcgen05).
 * Replaces nn.Linear forward (modeling_t5.py:305,310,528-536,581,1714; vit.py:17-20,41,53) and its autograd
 * dgrad/wgrad.  a_mn_major: A is stored [K][M] (lda = row stride) instead of [M][K]; b_mn_major: B is stored
 * [K][N] instead of [N][K].  Epilogue order: *alpha, +bias[N], (pre_out=copy), act, +residual, store.
 *   act: 0 none | 1 relu | 2 gelu(erf) | 3 multiply by relu'(aux) | 4 multiply by gelu'(aux) | 5, 6 fused cross entropy (below)
 *   atomic=1 (fp32 out only): atomicAdd into out; required when splits>1 (split-K over `splits` CTAs).
 * Dropout everywhere in this ABI is counter-based: p16 = round(p*65536), element kept iff a 16-bit hash of
 * (seed, element index) >= p16, kept values scaled by 65536/(65536-p16); p16 = 0 disables it. */
typedef struct vc_gemm_args {
  const void* A; const void* B;
  int64_t lda, ldb;
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  void* out; int64_t ldo; int32_t out_fp32; int32_t atomic;
  const float* bias;
  const float* residual; int64_t ldr;
  int32_t act;
  void* pre_out;
  const void* aux; int64_t ld_aux;
  float alpha;
  const float* alpha_dev;   /* optional device scalar multiplied into alpha (upstream loss gradient) */
  int32_t splits;
  int32_t tile_n;   /* 0 = auto, else 64/128/256 */
  uint32_t drop_seed, drop_p16;  /* dropout after act, before +residual; keep iff rnd16(seed, row*N+col) >= p16 */
  /* ---- LM head fused with the label-smoothed cross entropy (modeling_t5.py:1714-1721): the fp32 logits never exist.
   * act 5 (statistics pass): nothing is stored to `out`; every 128-row x tile_n block contributes, per row and per
   *   64*(tile_n/128)-column half block, a partial (max, sum exp(z - max), sum z) to ce_stats[row][slot][3] with
   *   slot = 2*(col/tile_n) + half, and the logit of the row's label to ce_zy[row].  tile_n must be given (128 or 256).
   * act 6 (gradient pass): out (bf16) = d loss / d logits = (exp(z - ce_lse[row]) - s/V - (1-s)[col == label]) / *ce_nvalid,
   *   zero rows for label -100.  vc_ce_combine turns the partials into ce_lse and the loss in between. */
  const int64_t* ce_labels; float* ce_stats; float* ce_zy; const float* ce_lse; const float* ce_nvalid; float ce_smoothing;
} vc_gemm_args;
int vc_gemm_bf16(const vc_gemm_args* args, void* stream);

/* ---- Fused attention forward (head_dim 64), softmax(scale*q.k^T + bias + mask).v without materialising scores.
 * Replaces modeling_t5.py:539-580 (T5Attention: unscaled scores + relative position bias + additive finfo.min key
 * mask, fp32 softmax) and vit.py:47-51 (scale 64^-0.5, no mask).  q/k/v are bf16 matrices [B*L, ld*] whose head h
 * occupies columns [*_col + 64h, +64).  bias_rel[h][k - q + Lq - 1] is the additive bias by relative position (the
 * T5 bucket table expanded by vc_bias_expand; NULL = zero bias, as in cross-attention modeling_t5.py:544-547).
 * kmask[b][k] = 1 to attend (HF get_extended_attention_mask semantics); causal adds k<=q (decoder self-attention).
 * out: bf16 [B*Lq, ldo], head h at columns [64h, 64h+64).  lse2: [B,H,Lq] log2-domain log-sum-exp (saved for bwd). */
typedef struct vc_attn_args {
  const void* q; const void* k; const void* v;
  int64_t ldq, ldk, ldv;
  int32_t q_col, k_col, v_col;
  int32_t B, H, Lq, Lk, head_dim;
  void* out; int64_t ldo;
  float* lse2;
  const float* bias_rel;
  const uint8_t* kmask;
  int32_t causal;
  float scale;
  uint32_t drop_seed, drop_p16;  /* dropout on the probabilities (modeling_t5.py:572-574, vit.py:49); index ((b*H+h)*Lq+q)*Lk+k */
  /* incremental decoding (forward only; leave 0 for training): queries sit at positions q + q_offset (+ *q_offset_dev),
   * K/V batches are kv_batch_rows rows apart (a KV cache of that capacity), bias row = bias_len entries with relative
   * position 0 at bias_zero (bias_len = 0: the training layout Lq+Lk-1 / Lq-1).  modeling_t5.py:484-488,500-525,551-556. */
  int32_t q_offset; const int32_t* q_offset_dev; int32_t kv_batch_rows; int32_t bias_zero, bias_len;
  /* self-attention over a padded sequence (Lq == Lk, the queries carry the key mask): query rows past the last attended
   * key are padding — no consumer can observe them (as keys they get an exactly-zero probability everywhere) — so whole
   * 128-query tiles of them are skipped: out rows = 0 (forward), no contribution (backward).  0 = compute every row. */
  int32_t q_like_k;
  /* incremental decoding with beam search (Lq == 1 only): sequence b reads the K/V of batch b / kv_batch_div, so the
   * num_beams hypotheses of one video share ONE copy of the cross-attention K/V (0 or 1: every sequence has its own). */
  int32_t kv_batch_div;
} vc_attn_args;
int vc_attn_fwd(const vc_attn_args* args, void* stream);

/* ---- Fused attention backward (autograd of the block above).  `fwd` repeats the forward arguments (fwd.out = O and
 * fwd.lse2 as saved by the forward).  dout: bf16 dL/dO [B*Lq, ld_do], head h at cols do_col + 64h.  delta: scratch
 * [B,H,Lq].  dq_acc: fp32 [B*Lq, ld_dq], cleared by the call then accumulated with atomics (head h at cols 64h).  dk/dv: bf16
 * [B*Lk, ld], head h at cols d*_col + 64h (fully written).  dbias_rel: fp32 [H, Lq+Lk-1] accumulated with atomics
 * (NULL when the bias is not a parameter); bucket_lut [Lq+Lk-1] lets tiles inside one bucket take a fast path. */
typedef struct vc_attn_bwd_args {
  vc_attn_args fwd;
  const void* dout; int64_t ld_do; int32_t do_col;
  float* delta;
  float* dq_acc; int64_t ld_dq;
  void* dk; int64_t ld_dk; int32_t dk_col;
  void* dv; int64_t ld_dv; int32_t dv_col;
  float* dbias_rel;
  const int32_t* bucket_lut;
} vc_attn_bwd_args;
int vc_attn_bwd(const vc_attn_bwd_args* args, void* stream);

/* ---- Normalisation.  kind 0: T5LayerNorm / RMS (modeling_t5.py:254-277, eps 1e-6, no bias); kind 1: nn.LayerNorm
 * (vit.py:64,70,96, eps 1e-5).  x fp32 [M,D]; y = (xhat*w (+bias)) * out_scale written as bf16 (out_bf16) and/or fp32
 * (out_f32) at row (r/rows_per_batch)*out_batch_stride + out_row_offset + r%rows_per_batch (rows_per_batch=0: row r).
 * rstd (and mean for kind 1) [M] are saved for the backward. */
int vc_norm_fwd(int kind, const float* x, const float* w, const float* bias, void* out_bf16, float* out_f32, float* rstd,
                float* mean, int M, int D, float eps, float out_scale, int rows_per_batch, int out_batch_stride,
                int out_row_offset, uint32_t drop_seed, uint32_t drop_p16, void* stream);
/* g = dL/dy (fp32, or bf16 when g_bf16), read through the same row map (and through the forward's output-dropout mask g_drop_*).
 * dx (+)= d/dx; dx_bf16 (optional) = bf16 copy of the final dx, masked by dxb_drop_* (the output dropout of the
 * sub-layer below, whose dY it is); dw/db accumulated with atomics (db only for kind 1; either may be NULL). */
int vc_norm_bwd(int kind, const void* g, int g_bf16 /* g is bf16 instead of fp32 */, const float* x, const float* w, const float* rstd, const float* mean, float* dx,
                void* dx_bf16, int accumulate_dx, float* dw, float* db, int M, int D, float scale, int rows_per_batch,
                int g_batch_stride, int g_row_offset, uint32_t g_drop_seed, uint32_t g_drop_p16, uint32_t dxb_drop_seed,
                uint32_t dxb_drop_p16, void* stream);

/* ---- Embedding (vid2seq.py:71, modeling_t5.py:972) and its scatter-add backward into the tied table. */
int vc_embed_fwd(const int64_t* ids, const float* table, float* out, int n, int d, int V, uint32_t drop_seed,
                 uint32_t drop_p16, void* stream);   /* + T5Stack input dropout (modeling_t5.py:1019), index i*d+c */
int vc_embed_bwd(const int64_t* ids, const float* dout, float* dtable, int n, int d, int V, uint32_t drop_seed,
                 uint32_t drop_p16, void* stream);
/* ---- labels = ids with pad->-100 (vid2seq.py:86-88); dec_in = shift_right(labels) (modeling_t5.py:845-868);
 * n_valid[0] = number of non-pad targets. */
int vc_prepare_targets(const int64_t* out_ids, int64_t* dec_in, int64_t* labels, float* n_valid, int B, int S,
                       int64_t pad_id, void* stream);
/* ---- out[h][r] = table[lut[r]][h] (modeling_t5.py:445-460 with lut = _relative_position_bucket :397-443 built on the
 * host with the reference's exact ops) and its backward dtable[lut[r]][h] += drel[h][r]. */
int vc_bias_expand(const float* table, const int32_t* lut, float* out, int H, int R, void* stream);
int vc_bias_fold(const float* drel, const int32_t* lut, float* dtable, int H, int R, void* stream);
/* ---- x + pos_embed with nearest interpolation when T != P (vit.py:119-127), and d(pos_embed). */
int vc_add_pos(const float* x, const float* pos, float* out, int B, int T, int C, int P, uint32_t drop_seed,
               uint32_t drop_p16, void* stream);     /* + pos_drop (vit.py:126) */
int vc_add_pos_bwd(const float* dx, float* dpos, int B, int T, int C, int P, uint32_t drop_seed, uint32_t drop_p16,
                   void* stream);
/* ---- F.cross_entropy(ignore_index=-100, label_smoothing) (modeling_t5.py:1721): loss_out[0] = mean over valid rows;
 * dlogits (bf16, optional) = d loss / d logits for upstream gradient 1. */
int vc_cross_entropy(const float* logits, int64_t ld, const int64_t* labels, const float* n_valid, float smoothing,
                     float* loss_out, void* dlogits_bf16, int64_t ldd, int n, int V, void* stream);
/* Combines the partials of the act-5 GEMM: lse[row] = log sum_c exp(z_c); loss_out[0] = mean over rows with label != -100 of
 * (1-s)*(lse - z_label) + s*(lse - mean_c z_c)  (F.cross_entropy(ignore_index=-100, label_smoothing=s)). */
int vc_ce_combine(const float* ce_stats, int n_slots, const float* ce_zy, const int64_t* labels, const float* n_valid,
                  float smoothing, int V, float* lse_out, float* loss_out, int M, void* stream);
/* ---- helpers: column sums (bias gradients), strided fp32->bf16 cast, row-block copy into the [video;text] memory. */
int vc_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, void* stream);
int vc_cast_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int M, int N, float scale, void* stream);
int vc_copy_rows_bf16(const void* src, void* dst, int B, int T, int C, int E, int row_off, void* stream);

/* ---- Incremental greedy decoding (model/vid2seq.py:150-162 with num_beams=1; HF-4.28 greedy semantics).
 * The step position lives in device memory (pos_dev) so one decode step is a fixed CUDA graph. */
int vc_kv_append(const void* src, int64_t lds, void* cache, int B, int cap, int C, const int32_t* pos_dev, void* stream);
int vc_greedy_next(const float* logits, int64_t ld, int V, uint8_t* done, int64_t* ids_out, int64_t* seq, int seq_ld,
                   const int32_t* pos_dev, int64_t eos_id, int64_t pad_id, int B, void* stream);
int vc_step_advance(int32_t* pos_dev, void* stream);
/* One linear layer of a decode step, M = batch rows (any M; 64 rows per CTA): out[M,N] = epi(A[M,K] . W[N,K]^T), W bf16.
 * A: bf16 [M][lda] (a_fp32 = 0) or the fp32 residual stream (a_fp32 = 1), optionally through the T5 RMS norm
 * (norm_w != NULL: x * rsqrt(mean x^2 + eps) * norm_w * out_scale, K <= 1024) — T5LayerNorm + nn.Linear in ONE launch
 * (modeling_t5.py:254-277 + :305,310,528-536,581).  Epilogue: relu, + residual (fp32 out, may alias), bf16 / fp32 store. */
int vc_decode_linear(const void* A, int64_t lda, int a_fp32, const float* norm_w, float eps, float out_scale, const void* W,
                     int64_t ldw, void* out, int64_t ldo, int out_fp32, const float* residual, int64_t ldr, int relu, int M,
                     int N, int K, void* stream);
/* Beam search (vid2seq.py:150-162 with num_beams > 1; HF-4.28 GenerationMixin.beam_search, third-party): for every batch
 * item the 2*num_beams best  log_softmax(logits[b*num_beams + r])[v] + beam_scores[b*num_beams + r]  over (r, v), sorted
 * descending -> out_scores / out_tokens (v) / out_beams (r), each [B, 2*num_beams].  num_beams <= 8. */
int vc_beam_topk(const float* logits, int64_t ld, int V, const float* beam_scores, int num_beams, int B, float* out_scores,
                 int32_t* out_tokens, int32_t* out_beams, void* stream);
/* KV-cache rows [0, n) of every sequence follow their beam: dst[b] = src[beam_idx[b]] (HF _reorder_cache,
 * modeling_t5.py:1771-1793).  src/dst: bf16 [Bn, cap, C], distinct buffers. */
int vc_kv_reorder(const void* src, void* dst, const int32_t* beam_idx, int Bn, int cap, int C, int n, void* stream);

/* ---- Optimiser tail over the flat parameter buffer (dvc.py:112-126). */
int vc_sumsq(const float* g, int64_t n, float* out_accum, void* stream);               /* out_accum[0] += |g|^2 */
/* clip_grad_norm_ (coef from *norm_sq, skipped if clip_max_norm<=0 or norm_sq NULL) + torch.optim.Adam step +
 * bf16 shadow re-pack, one pass.  grad_scale multiplies g first (1/world_size for data parallel averaging). */
int vc_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                 float eps, int step, const float* norm_sq, float clip_max_norm, float grad_scale, void* stream);
/* rows [V-num_bins, V) /= mean||rows|| / mean||rows [0, V-num_bins)||  (dvc.py:118-126); scratch2 = 2 floats. */
int vc_renorm_time_tokens(float* w, void* w_bf16, int V, int d, int num_bins, float* scratch2, void* stream);
int vc_cast_flat_bf16(const float* src, void* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDCHAP_H_ */
