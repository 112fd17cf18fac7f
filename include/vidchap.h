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

/* ---- GEMM: out[M,N] = epilogue( alpha * op(A)[M,K] . op(B)[N,K]^T ), bf16 operands, fp32 accumulate (tcgen05).
 * Replaces nn.Linear forward (modeling_t5.py:305,310,528-536,581,1714; vit.py:17-20,41,53) and its autograd
 * dgrad/wgrad.  a_mn_major: A is stored [K][M] (lda = row stride) instead of [M][K]; b_mn_major: B is stored
 * [K][N] instead of [N][K].  Epilogue order: *alpha, +bias[N], (pre_out=copy), act, +residual, store.
 *   act: 0 none | 1 relu | 2 gelu(erf) | 3 multiply by relu'(aux) | 4 multiply by gelu'(aux)
 *   atomic=1 (fp32 out only): atomicAdd into out; required when splits>1 (split-K over `splits` CTAs). */
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
  int32_t splits;
  int32_t tile_n;   /* 0 = auto, else 64/128/256 */
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
} vc_attn_args;
int vc_attn_fwd(const vc_attn_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDCHAP_H_ */
