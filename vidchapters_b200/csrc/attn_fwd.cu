// Fused attention forward on tcgen05 (sm_100a), head_dim 64.
//
// Replaces T5Attention.forward's score/softmax/PV chain (reference model/modeling_t5.py:539-580: UNSCALED q.k^T +
// position_bias (+ additive key mask), fp32 softmax, .v) and the ViT attention (model/vit.py:47-51: scale 1/8, no
// mask).  The (B,H,Lq,Lk) score / probability tensors of the reference never exist in HBM.
//
// One CTA = one (batch, head, 128-query tile); 2 CTAs per SM (112 KB smem, 256 TMEM columns each) so one CTA's
// softmax overlaps the other's MMAs.  320 threads:
//   warp 0 lane 0 : TMA producer (Q once; K,V 128-key tiles through a 2-stage ring), 3-D maps zero-fill past Lq/Lk
//   warp 1        : TMEM alloc; lane 0 issues  S = Q.K^T (128x128x64)  and  O += P.V (128x64x128, accumulating in TMEM)
//   warps 2..9    : softmax, two threads per query row (64 keys of the tile each).  Pass 1 row max, pass 2
//                   p = exp2(s2 - m) -> bf16 P tile written into 128B-swizzled smem (the A operand of the PV MMA);
//                   online softmax with an integer running max; O is rescaled in TMEM only when that max grows.
// Scores are handled in the log2 domain: s2 = (acc*scale + bias) * log2(e).  Masked keys (key-padding or causal)
// take the reference's additive finfo.min semantics (a fully masked row degenerates to uniform, like the reference);
// columns past Lk are excluded exactly.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

constexpr int kTQ = 128, kTK = 128, kD = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMasked = -3.0e38f;  // stands for the reference's (1-mask)*finfo(float32).min after the add

struct AttnFwdParams {
  int B, H, Lq, Lk;
  int q_col, k_col, v_col;
  __nv_bfloat16* out;
  long long ldo;
  float* lse2;            // [B,H,Lq] log2-domain logsumexp of s2
  const float* bias_rel;  // [H][Lq+Lk-1] or null; index (k - q + Lq - 1)
  const uint8_t* kmask;   // [B][Lk] or null
  int causal;
  float scale_log2e;
  uint32_t drop_seed, drop_p16;
  const uint32_t* drop_salt;
  int q_offset;              // absolute position of query row 0 (incremental decoding), added to *q_offset_dev
  const int* q_offset_dev;   // optional device scalar (CUDA-graph friendly decode step counter)
  int bias_zero, bias_len;   // bias row: index of relative position 0, row length
};

// smem: Q 16K | K 2x16K | V 16K | P 32K | bias window (Lk+128 floats) | key mask (Lk bytes) | barriers.
// The bias/mask staging matters: with two 100 KB CTAs per SM almost no L1 is left, so per-element global loads of the
// bias row went to L2 and made the kernel 10x slower (profiles/r01_launches_before.txt).
constexpr int kAttnFwdTiles = 16384 + 2 * 16384 + 16384 + 32768;
constexpr int kAttnMaxLk = 1536;  // bias window + key ceilings + row-statistic exchange must fit next to 96 KB of tiles, twice per SM
constexpr int kAttnFwdTail = 256 + 2048 + 1024;   // barriers + flags | sMx | sL

__global__ void __launch_bounds__(320, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = sK + 2 * 16384;
  uint8_t* sP = sV + 16384;
  float* sBias = reinterpret_cast<float*>(sP + 32768);
  const int lk_pad = ((p.Lk + kTK - 1) / kTK) * kTK;   // key count rounded up to whole tiles
  // per-key ceiling applied with one FMNMX: +inf attend | kMasked (reference's additive finfo.min) | -inf out of range
  float* sPen = sBias + lk_pad + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPen + lk_pad);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  // Fast-path flags per 128-slot block: the bias changes inside the block / some key of the block is masked or out of
  // range.  A key tile whose bias window is one constant and whose keys all attend (T5: every tile further than 128
  // positions from the diagonal; cross-attention and the ViT: every full tile) needs no per-element bias / ceiling
  // lookup: s2 = acc*scale + c.
  int* sFlagB = reinterpret_cast<int*>(bars + 12);   // [16]
  int* sFlagP = sFlagB + 16;                          // [16]
  int* sLastKey = sFlagP + 16;                        // index of the last key that attends (-1: none)
  float* sMx = reinterpret_cast<float*>(bars + 32);   // [2 parities][2 halves][128 rows] half-row tile maxima
  float* sL = sMx + 512;                              // [2 halves][128 rows] half-row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kTQ;
  const int qoff = p.q_offset + (p.q_offset_dev ? __ldg(p.q_offset_dev) : 0);  // queries sit at positions q + qoff

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);     // one arrival per softmax warp (each arrival wakes every thread sleeping on a barrier)
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) sFlagB[threadIdx.x] = 0;   // (covers sFlagP too)
  if (threadIdx.x == 32) *sLastKey = -1;
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  __syncthreads();
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is read from here on
  pdl_trigger();
  // stage this query tile's bias window (index k + 127 - r == (k - q + Lq - 1) - (Lq - 128 - q0)) and the key mask
  {
    const int win0 = p.bias_zero - 127 - q0 - qoff;  // global bias index of window slot 0 (may be negative: never used)
    // Both arrays are always filled (zeros / ones without bias / mask) and padded to whole tiles so that the
    // softmax loops below are branch-free: conditional loads compiled to one divergence region per element and
    // serialised the whole tile (profiles/r01_attn_notes.md).
    const float* brow = p.bias_rel ? p.bias_rel + (long long)h * p.bias_len : nullptr;
    for (int i = threadIdx.x; i < lk_pad + 128; i += blockDim.x) {
      const int gi = win0 + i;
      const float val = (brow && gi >= 0 && gi < p.bias_len) ? __ldg(brow + gi) * kLog2e : 0.f;
      sBias[i] = val;
      if (brow && i > 0) {
        const float prev = (gi - 1 >= 0 && gi - 1 < p.bias_len) ? __ldg(brow + gi - 1) * kLog2e : 0.f;
        if (val != prev) sFlagB[i >> 7] = 1;
      }
    }
    const uint8_t* mrow = p.kmask ? p.kmask + (long long)b * p.Lk : nullptr;
    for (int i = threadIdx.x; i < lk_pad; i += blockDim.x) {
      const float pen = (i >= p.Lk) ? -INFINITY : ((mrow && mrow[i] == 0) ? kMasked : INFINITY);
      sPen[i] = pen;
      if (pen != INFINITY) sFlagP[i >> 7] = 1;
      else atomicMax(sLastKey, i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns

  // number of key tiles this query tile visits (causal: skip tiles entirely above the diagonal)
  int nkt = (p.Lk + kTK - 1) / kTK;
  if (p.causal) nkt = min(nkt, (q0 + qoff + kTQ - 1) / kTK + 1);
  // Trailing key tiles made only of masked (padding) keys contribute exp2(finfo.min - m) == 0 exactly to every row as
  // long as the row attends to at least one key (key 0 of a padded sequence always does; causal rows see key 0), so
  // they are skipped.  With no attended key at all the reference degenerates to a uniform softmax: keep every tile.
  if (*sLastKey >= 0 && (!p.causal || sPen[0] == INFINITY)) nkt = min(nkt, *sLastKey / kTK + 1);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_arrive_expect_tx(q_full, 16384);
    tma_load_3d(sQ, &tmQ, q_full, p.q_col + h * kD, q0, b);
    for (int j = 0; j < nkt; ++j) {
      const int st = j & 1;
      mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx(&k_full[st], 16384);
      tma_load_3d(sK + st * 16384, &tmK, &k_full[st], p.k_col + h * kD, j * kTK, b);
      mbar_wait(v_empty, (j & 1) ^ 1);
      mbar_arrive_expect_tx(v_full, 16384);
      tma_load_3d(sV, &tmV, v_full, p.v_col + h * kD, j * kTK, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S: A=Q K-major, B=K K-major, N=128
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O: A=P K-major, B=V MN-major, N=64
    mbar_wait(q_full, 0);
    const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ), 0, 1024);
    for (int j = 0; j < nkt; ++j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + st * 16384), 0, 1024);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k) tc_mma_bf16(tmem_S, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k > 0);
      tc_commit(&k_empty[st]);  // K stage is free as soon as S(j) has been computed
      tc_commit(s_full);
      // P(j) written and S(j) consumed; O(j-1) read back; V(j) landed
      mbar_wait(p_full, j & 1);   // (also covers any in-place rescale of O by the softmax warps)
      mbar_wait(v_full, j & 1);
      tc_fence_after();
      const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV), 0, 1024);
#pragma unroll
      for (int k = 0; k < kTK / 16; ++k) {
        const uint64_t pd = make_smem_desc_sw128(smem_u32(sP + (k >> 2) * 16384) + (k & 3) * 32, 0, 1024);
        tc_mma_bf16(tmem_O, pd, vdesc + (uint64_t)(k * 128), idesc_o, (j > 0 || k > 0));   // O += P(j).V(j)
      }
      tc_commit(v_empty);
      if (j == nkt - 1) tc_commit(o_full);
    }
  } else if (warp >= 2) {
    // ===================== softmax / epilogue: 8 warps, two threads per query row =====================
    // Thread (r, hf) owns keys [64 hf, 64 hf + 64) of every tile and output columns [32 hf, 32 hf + 32) of row r (TMEM
    // lane r: warp w may only touch lanes 32 (w % 4) ..).  The two halves of a row exchange their tile maxima through
    // shared memory; the running sum stays split until the end.  O accumulates in TMEM across tiles (the PV MMA adds
    // into it); it is rescaled in place only when a row's INTEGER running maximum grows — after the first tiles almost
    // never — instead of being read back and re-accumulated in registers every tile.
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;  // row in tile == TMEM lane
    const int q = q0 + r;            // row inside this call's query block (output / lse index)
    const int q_abs = q + qoff;      // its sequence position (bias window, causal mask)
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float* brow = sBias + (127 - r);  // brow[k] = bias(k - q) * log2e
    const uint32_t drop_rk = drop_row_key(drop_salted(p.drop_seed, p.drop_salt), ((unsigned long long)b * p.H + h) * p.Lq + q);
    float m_run = -INFINITY, l_run = 0.0f;

    for (int j = 0; j < nkt; ++j) {
      const int k0 = j * kTK;
      const int kh = k0 + hf * 64;   // first key of this thread's half of the tile
      mbar_wait(s_full, j & 1);      // (tensor pipe is in order: S(j) done implies PV(j-1) done — sP and O are free)
      tc_fence_after();
      // ---- pass 1: row max of s2 over the tile.  s2 = min(acc*scale*log2e + bias*log2e, pen[k]); causal tiles add k<=q.
      const bool causal_tile = p.causal && (k0 + kTK - 1 > q0 + qoff);   // uniform: only tiles touching the diagonal
      // uniform over the CTA: the tile's whole bias window is one value and every key attends
      const bool fast = !causal_tile && (sFlagB[j] | sFlagB[j + 1] | sFlagP[j]) == 0;
      const float cb = sBias[k0 + 127];
      float m_loc = -INFINITY;
      if (fast) {
        float v[32], w[32];
        tmem_ld32(tmem_S + lane_off + hf * 64, v);
        tmem_ld32(tmem_S + lane_off + hf * 64 + 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m_loc = fmaxf(m_loc, fmaxf(v[i], w[i]));
        m_loc = fmaf(m_loc, p.scale_log2e, cb);   // scale > 0: max commutes with the affine map
      } else {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(tmem_S + lane_off + hf * 64 + c * 32, v);
          const float4* pen4 = reinterpret_cast<const float4*>(sPen + kh + c * 32);
          const float* bk = brow + kh + c * 32;
          const int tq = q_abs - kh - c * 32;  // column i is causally masked iff i > tq
          tmem_ld_wait();
          // s2 (+ causal) is written back to TMEM so that pass 2 only has to subtract the max and exponentiate
          if (causal_tile) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 pe = pen4[i >> 2];
              const float pen[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float s2 = fminf(fmaf(v[i + e], p.scale_log2e, bk[i + e]), pen[e]);
                s2 = (i + e > tq) ? fminf(s2, kMasked) : s2;
                v[i + e] = s2;
                m_loc = fmaxf(m_loc, s2);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 pe = pen4[i >> 2];
              const float pen[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[i + e] = fminf(fmaf(v[i + e], p.scale_log2e, bk[i + e]), pen[e]);
                m_loc = fmaxf(m_loc, v[i + e]);
              }
            }
          }
          tmem_st32(tmem_S + lane_off + hf * 64 + c * 32, v);
        }
        tmem_st_wait();
      }
      // exchange the half-row maxima (slots double-buffered by tile parity)
      float* mx = sMx + (j & 1) * 256;
      mx[hf * 128 + r] = m_loc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // Integer running max (log2 domain): every rescale factor is an exact power of two, so the bf16 rounding of
      // P = 2^(s2 - m) does not depend on the tiling / on when the maximum was discovered.
      const float m_new = fmaxf(m_run, ceilf(fmaxf(m_loc, mx[(hf ^ 1) * 128 + r])));
      const float corr = fast_exp2(m_run - m_new);  // first tile: m_run = -inf -> 0
      if (j > 0 && __any_sync(0xffffffffu, m_new > m_run)) {   // rare: rescale this warp's 32 rows x 32 columns of O
        float o[32];
        tmem_ld32(tmem_O + lane_off + hf * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= corr;
        tmem_st32(tmem_O + lane_off + hf * 32, o);
        tmem_st_wait();
      }
      // ---- pass 2: p = exp2(s2 - m_new), write bf16 P (swizzled, K-major A operand), row sum
      float l_tile = 0.0f;
      // fast tiles still hold the raw accumulator: one FMA folds scale, bias and the max; slow tiles hold s2
      const float e_mul = fast ? p.scale_log2e : 1.0f;
      const float e_add = fast ? cb - m_new : -m_new;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(tmem_S + lane_off + hf * 64 + c * 32, v);
        tmem_ld_wait();
        // the normaliser l uses the un-dropped probabilities (dropout acts on softmax's output)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float pv = fast_exp2(fmaf(v[i], e_mul, e_add));
          l_tile += pv;
          v[i] = pv;
        }
        if (p.drop_p16) {
          // mask row = (b, h, q), column = k.  Kept probabilities stay unscaled here: O = (sum mask.p.v) * sc / l, the
          // factor sc = 1/(1-p) is folded into the final normalisation.
          drop_select<32>(v, drop_rk, p.drop_p16, (uint32_t)(kh + c * 32));
        }
        // 32 columns = 4 x 16-byte chunks of this row; the 64-wide swizzle atom is the thread's half hf
        uint8_t* prow = sP + hf * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = (c * 4 + g) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + ch * 16) =
              make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                         pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      l_run = l_run * corr + l_tile;
      m_run = m_new;
    }
    // ---- epilogue: combine the two half-row sums, normalise O (TMEM) and store
    sL[hf * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_row = l_run + sL[(hf ^ 1) * 128 + r];
    mbar_wait(o_full, 0);
    tc_fence_after();
    {
      float o[32];
      tmem_ld32(tmem_O + lane_off + hf * 32, o);
      tmem_ld_wait();
      if (q < p.Lq) {
        const float inv = (p.drop_p16 ? drop_scale(p.drop_p16) : 1.0f) / l_row;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.Lq + q) * p.ldo + h * kD + hf * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          dst[g] = make_uint4(pack_bf16x2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv),
                              pack_bf16x2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv),
                              pack_bf16x2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv),
                              pack_bf16x2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv));
        if (p.lse2 && hf == 0) p.lse2[((long long)b * p.H + h) * p.Lq + q] = m_run + log2f(l_row);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace vc

using namespace vc;

namespace vc {
int launch_attn_fwd_pair(const vc_attn_args* a, cudaStream_t st);   // attn_fwd2.cu
int launch_attn_decode(const vc_attn_args* a, cudaStream_t st);     // decode2.cu
}

extern "C" int vc_attn_fwd(const vc_attn_args* a, void* stream) {
  VC_CHECK(a != nullptr, "vc_attn_fwd: null args");
  VC_CHECK(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "vc_attn_fwd: bad dims");
  VC_CHECK(a->head_dim == 64, "vc_attn_fwd: head_dim must be 64 (got %d)", a->head_dim);
  VC_CHECK(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "vc_attn_fwd: strides must be x8");
  VC_CHECK(a->q_col % 8 == 0 && a->k_col % 8 == 0 && a->v_col % 8 == 0, "vc_attn_fwd: column offsets must be x8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VC_CHECK(a->scale > 0.f, "vc_attn_fwd: scale must be positive");
  VC_CHECK(!a->q_like_k || (a->Lq == a->Lk && a->kmask), "vc_attn_fwd: q_like_k needs self-attention with a key mask");
  // one query per sequence (incremental decoding): stream the KV cache (decode2.cu); VIDCHAP_ATTN_DECODE=0 = A/B switch
  static const bool dec_on = [] { const char* e = getenv("VIDCHAP_ATTN_DECODE"); return !(e && e[0] == '0'); }();
  if (dec_on && a->Lq == 1 && a->drop_p16 == 0 && a->lse2 == nullptr && !a->q_like_k) return launch_attn_decode(a, st);
  VC_CHECK(a->kv_batch_div <= 1, "vc_attn_fwd: kv_batch_div needs the single-query decode kernel");
  // training shapes: two query tiles per CTA sharing one K/V stream (attn_fwd2.cu); VIDCHAP_ATTN_FWD_PAIR=0 = A/B switch
  static const bool pair_on = [] { const char* e = getenv("VIDCHAP_ATTN_FWD_PAIR"); return !(e && e[0] == '0'); }();
  if (pair_on && a->Lq > 128 && a->q_offset == 0 && a->q_offset_dev == nullptr && a->kv_batch_rows == 0 && a->bias_len == 0)
    return launch_attn_fwd_pair(a, st);
  CUtensorMap tmQ, tmK, tmV;
  int s;
  const uint64_t kv_rows = a->kv_batch_rows > 0 ? a->kv_batch_rows : a->Lk;  // rows between batches of K/V in memory
  VC_CHECK(kv_rows >= (uint64_t)a->Lk, "vc_attn_fwd: kv_batch_rows < Lk");
  if ((s = make_tmap_3d(&tmQ, a->q, a->ldq, a->Lq, a->B, a->ldq, (uint64_t)a->Lq * a->ldq, 64, kTQ)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmK, a->k, a->ldk, a->Lk, a->B, a->ldk, kv_rows * a->ldk, 64, kTK)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmV, a->v, a->ldv, a->Lk, a->B, a->ldv, kv_rows * a->ldv, 64, kTK)) != VC_OK) return s;
  AttnFwdParams p;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.q_col = a->q_col; p.k_col = a->k_col; p.v_col = a->v_col;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  p.lse2 = a->lse2; p.bias_rel = a->bias_rel; p.kmask = a->kmask; p.causal = a->causal;
  p.scale_log2e = a->scale * kLog2e;
  p.drop_seed = a->drop_seed; p.drop_p16 = a->drop_p16; p.drop_salt = drop_salt_ptr();
  p.q_offset = a->q_offset; p.q_offset_dev = a->q_offset_dev;
  p.bias_zero = a->bias_len > 0 ? a->bias_zero : a->Lq - 1;
  p.bias_len = a->bias_len > 0 ? a->bias_len : a->Lq + a->Lk - 1;
  VC_CHECK(a->Lk <= kAttnMaxLk, "vc_attn_fwd: Lk=%d exceeds the %d keys the bias/mask staging supports", a->Lk, kAttnMaxLk);
  const int lk_pad = ((a->Lk + kTK - 1) / kTK) * kTK;
  VC_CHECK(a->scale > 0.f, "vc_attn_fwd: scale must be positive");
  const int smem_bytes = kAttnFwdTiles + (lk_pad + 128) * 4 + lk_pad * 4 + kAttnFwdTail;
  static PerDeviceOnce attr;
  if (attr.need()) {
    VC_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kAttnFwdTiles + (kAttnMaxLk + 128) * 4 + kAttnMaxLk * 4 + kAttnFwdTail));
  }
  dim3 grid((a->Lq + kTQ - 1) / kTQ, a->H, a->B);
  VC_CUDA(launch_kernel(attn_fwd_kernel, grid, dim3(320), (size_t)smem_bytes, st, tmQ, tmK, tmV, p));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
