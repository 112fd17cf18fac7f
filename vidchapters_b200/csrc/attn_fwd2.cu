// Fused attention forward, pair variant (sm_100a, head_dim 64): one CTA = one (batch, head) and TWO 128-query tiles.
//
// Same arithmetic and rounding points as attn_fwd.cu (reference: model/modeling_t5.py:539-580, model/vit.py:47-51) —
// what changes is the schedule.  The single-tile kernel alternates  S = Q.K^T  ->  softmax  ->  O += P.V  per CTA and
// reads every score tile from tensor memory two or three times; ncu showed the tensor pipe 13 % busy with the softmax
// warps stalled on TMEM loads and on the tile's MMAs (profiles/r01_ncu_prof_attn_fwd_r01f.txt).  Here
//   * a group releases its S buffer as soon as it has read the tile's last chunk in the exp pass (TMEM reads are cheap:
//     ~900 B/clk/SM measured, tools/microbench), so the MMA warp issues S(j+1) of that group while the group is still
//     exponentiating, packing and storing tile j, and PV(j) no longer sits between a group and its next scores;
//   * two softmax groups (one per query tile, 8 warps each, two threads per row) share one K/V stream (3-stage rings,
//     each K/V tile is fetched once for 256 queries instead of once per 128) and ping-pong on the tensor pipe, issue
//     order  S_A(j+1), PV_A(j), S_B(j+1), PV_B(j);
//   * the O accumulators stay in TMEM across key tiles (rescaled in place only when a row's integer running maximum
//     grows), P goes through 128B-swizzled shared memory as the A operand of the PV MMA.
// TMEM: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).  576 threads, 1 CTA / SM (211 KB of shared memory):
//   warp 0 lane 0 : TMA producer          warp 1 : TMEM alloc, lane 0 issues the MMAs
//   warps 2..9    : softmax group A       warps 10..17 : softmax group B
// Query tiles made only of padding (self-attention, `q_like_k`: the query sequence carries the key mask) are skipped:
// their rows can never reach the loss — as keys they are masked with an exactly-zero probability in every consumer — so
// the group stores zeros for them.  vc_attn_fwd selects this kernel for Lq > 128 (training shapes); the incremental
// decoding fields stay with attn_fwd.cu.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

constexpr int kP2TQ = 128, kP2TK = 128, kP2D = 64;
constexpr int kP2Stages = 3;
constexpr float kP2Log2e = 1.4426950408889634f;
constexpr float kP2Masked = -3.0e38f;
constexpr int kP2MaxLk = 1536;

struct AttnFwd2Params {
  int B, H, Lq, Lk;
  int q_col, k_col, v_col;
  __nv_bfloat16* out;
  long long ldo;
  float* lse2;
  const float* bias_rel;   // [H][Lq+Lk-1] or null; index (k - q + Lq - 1)
  const uint8_t* kmask;    // [B][Lk] or null
  int causal;
  int q_like_k;            // queries past the last attended key are padding: skip their tiles
  float scale_log2e;
  uint32_t drop_seed, drop_p16;
  const uint32_t* drop_salt;
  long long* trace;        // debug (vc_debug_set_trace): clock64 timeline of CTA (0,0,0)
};

#define VC_TRACE2(slot, ev)                                                                                  \
  do {                                                                                                       \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (slot) < 2048)                    \
      p.trace[(slot)] = ((long long)(ev) << 48) | (clock64() & 0xFFFFFFFFFFFFLL);                             \
  } while (0)

constexpr int kP2Tiles = 2 * 16384 + kP2Stages * 16384 * 2 + 2 * 32768;                 // Q | K ring | V ring | P x2
constexpr int kP2Tail = 512 /*barriers + flags*/ + 4096 /*sMx*/ + 2048 /*sL*/;
static inline int attn_fwd2_smem(int lk_pad) { return kP2Tiles + (lk_pad + 256) * 4 + lk_pad * 4 + kP2Tail; }

__global__ void __launch_bounds__(576, 1)
attn_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnFwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [2][16 KB]
  uint8_t* sK = sQ + 2 * 16384;                         // [3][16 KB]
  uint8_t* sV = sK + kP2Stages * 16384;                 // [3][16 KB]
  uint8_t* sP = sV + kP2Stages * 16384;                 // [2 groups][32 KB]
  float* sBias = reinterpret_cast<float*>(sP + 2 * 32768);
  const int lk_pad = ((p.Lk + kP2TK - 1) / kP2TK) * kP2TK;
  float* sPen = sBias + lk_pad + 256;                   // per-key ceiling: +inf attend | kP2Masked | -inf out of range
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPen + lk_pad);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [3]
  uint64_t* k_empty = bars + 4;    // [3]
  uint64_t* v_full = bars + 7;     // [3]
  uint64_t* v_empty = bars + 10;   // [3]
  uint64_t* s_full = bars + 13;    // [2] MMA -> group: S_g(j) complete
  uint64_t* s_free = bars + 15;    // [2] group -> MMA: S_g(j) is in registers
  uint64_t* p_full = bars + 17;    // [2] group -> MMA: P_g(j) in shared memory (and O_g rescaled)
  uint64_t* pv_done = bars + 19;   // [2] MMA -> group: O_g += P_g(j).V(j) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  int* sFlagB = reinterpret_cast<int*>(bars + 22);      // [16] bias changes inside the 128-slot block
  int* sFlagP = sFlagB + 16;                             // [16] some key of the 128-key block is masked / out of range
  int* sLastKey = sFlagP + 16;                           // last key that attends (-1: none)
  float* sMx = reinterpret_cast<float*>(bars + 64);      // [2 groups][2 parities][2 halves][128]
  float* sL = sMx + 1024;                                // [2 groups][2 halves][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0p = qp * 2 * kP2TQ;    // first query of the pair

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kP2Stages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 8); mbar_init(&p_full[g], 8); mbar_init(&pv_done[g], 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) sFlagB[threadIdx.x] = 0;   // (covers sFlagP too)
  if (threadIdx.x == 32) *sLastKey = -1;
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  {
    // bias window of the 256 query rows: slot i <-> relative position (k - q) with i = k + 255 - (q - q0p)
    const int win0 = (p.Lq - 1) - 255 - q0p;     // bias index of slot 0 (may be negative: such slots are never read)
    const int blen = p.Lq + p.Lk - 1;
    const float* brow = p.bias_rel ? p.bias_rel + (long long)h * blen : nullptr;
    for (int i = threadIdx.x; i < lk_pad + 256; i += blockDim.x) {
      const int gi = win0 + i;
      const float val = (brow && gi >= 0 && gi < blen) ? __ldg(brow + gi) * kP2Log2e : 0.f;
      sBias[i] = val;
      if (brow && i > 0) {
        const float prev = (gi - 1 >= 0 && gi - 1 < blen) ? __ldg(brow + gi - 1) * kP2Log2e : 0.f;
        if (val != prev) sFlagB[i >> 7] = 1;
      }
    }
    const uint8_t* mrow = p.kmask ? p.kmask + (long long)b * p.Lk : nullptr;
    for (int i = threadIdx.x; i < lk_pad; i += blockDim.x) {
      const float pen = (i >= p.Lk) ? -INFINITY : ((mrow && mrow[i] == 0) ? kP2Masked : INFINITY);
      sPen[i] = pen;
      if (pen != INFINITY) sFlagP[i >> 7] = 1;
      else atomicMax(sLastKey, i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int last_key = *sLastKey;

  // key tiles visited by each group (uniform over the CTA)
  int nkt[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int q0 = q0p + g * kP2TQ;
    int n = (p.Lk + kP2TK - 1) / kP2TK;
    if (p.causal) n = min(n, (q0 + kP2TQ - 1) / kP2TK + 1);
    // trailing key tiles of padding contribute exactly 0 as long as every row attends somewhere (see attn_fwd.cu)
    if (last_key >= 0 && (!p.causal || sPen[0] == INFINITY)) n = min(n, last_key / kP2TK + 1);
    if (q0 >= p.Lq) n = 0;                                              // no such tile
    if (p.q_like_k && last_key >= 0 && q0 > last_key) n = 0;            // queries of this tile are all padding
    nkt[g] = n;
  }
  const int nkt_max = max(nkt[0], nkt[1]);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    if (nkt_max > 0) {
      mbar_arrive_expect_tx(q_full, 2 * 16384);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col + h * kP2D, q0p, b);
      tma_load_3d(sQ + 16384, &tmQ, q_full, p.q_col + h * kP2D, q0p + kP2TQ, b);   // rows past Lq arrive as zeros
      for (int j = 0; j < nkt_max; ++j) {
        const int st = j % kP2Stages;
        const uint32_t ph = (uint32_t)(j / kP2Stages) & 1u;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], 16384);
        tma_load_3d(sK + st * 16384, &tmK, &k_full[st], p.k_col + h * kP2D, j * kP2TK, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], 16384);
        tma_load_3d(sV + st * 16384, &tmV, &v_full[st], p.v_col + h * kP2D, j * kP2TK, b);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    if (nkt_max > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S: A = Q K-major, B = K K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O: A = P K-major, B = V MN-major
      mbar_wait(q_full, 0);
      auto issue_s = [&](int g, int j) {
        const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + g * 16384), 0, 1024);
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + (j % kP2Stages) * 16384), 0, 1024);
#pragma unroll
        for (int k = 0; k < kP2D / 16; ++k)
          tc_mma_bf16(tmem_base + g * 128, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k > 0);
        tc_commit(&s_full[g]);
      };
      // prologue: S_A(0), S_B(0)
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int g = 0; g < 2; ++g)
        if (nkt[g] > 0) issue_s(g, 0);
      tc_commit(&k_empty[0]);
      for (int j = 0; j < nkt_max; ++j) {
        const int st = j % kP2Stages;
        const uint32_t ph = (uint32_t)(j / kP2Stages) & 1u;
        const int stn = (j + 1) % kP2Stages;
        const uint32_t phn = (uint32_t)((j + 1) / kP2Stages) & 1u;
        bool k_ready = false, v_ready = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (j + 1 < nkt[g]) {            // S_g(j+1): the group has read S_g(j) out of tensor memory
            if (!k_ready) { mbar_wait(&k_full[stn], phn); k_ready = true; }
            VC_TRACE2(1024 + j * 16 + g * 8 + 0, 100 + g * 10);
            mbar_wait(&s_free[g], j & 1);
            tc_fence_after();
            VC_TRACE2(1024 + j * 16 + g * 8 + 1, 101 + g * 10);
            issue_s(g, j + 1);
            VC_TRACE2(1024 + j * 16 + g * 8 + 2, 102 + g * 10);
          }
          if (j < nkt[g]) {                // O_g += P_g(j).V(j)
            if (!v_ready) { mbar_wait(&v_full[st], ph); v_ready = true; }
            mbar_wait(&p_full[g], j & 1);
            tc_fence_after();
            VC_TRACE2(1024 + j * 16 + g * 8 + 3, 103 + g * 10);
            const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + st * 16384), 0, 1024);
#pragma unroll
            for (int k = 0; k < kP2TK / 16; ++k) {
              const uint64_t pd = make_smem_desc_sw128(smem_u32(sP + g * 32768 + (k >> 2) * 16384) + (k & 3) * 32, 0, 1024);
              tc_mma_bf16(tmem_base + 256 + g * 64, pd, vdesc + (uint64_t)(k * 128), idesc_o, (j > 0 || k > 0));
            }
            tc_commit(&pv_done[g]);
            VC_TRACE2(1024 + j * 16 + g * 8 + 4, 104 + g * 10);
          }
        }
        if (j + 1 < nkt_max) tc_commit(&k_empty[stn]);   // K(j+1) consumed by every S(j+1) issued above
        tc_commit(&v_empty[st]);                          // V(j) consumed
      }
    }
  } else if (warp >= 2) {
    // ===================== softmax / epilogue: group g = 8 warps, two threads per query row =====================
    const int g = (warp - 2) >> 3;                 // 0: warps 2..9, 1: warps 10..17
    const int wg = (warp - 2) & 7;                 // warp inside the group
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int hf = wg >> 2;                        // which 64 keys of the tile / 32 output columns ((quarter, hf) is unique per group)
    const int r = quarter * 32 + lane;             // row inside the tile == TMEM lane
    const int q0 = q0p + g * kP2TQ;
    const int q = q0 + r;
    const int n_g = nkt[g];
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off + hf * 64;
    const uint32_t tO = tmem_base + 256 + g * 64 + lane_off + hf * 32;
    const float* brow = sBias + (255 - (g * kP2TQ + r));   // brow[k] = bias(k - q) * log2e
    const uint32_t drop_rk = drop_row_key(drop_salted(p.drop_seed, p.drop_salt), ((unsigned long long)b * p.H + h) * p.Lq + q);
    float* mxg = sMx + g * 512;
    uint8_t* sPg = sP + g * 32768;
    const int bar_id = 1 + g;
    float m_run = -INFINITY, l_run = 0.0f;

    if (n_g == 0) {
      // padding-only (or absent) query tile: define the outputs (zeros) and leave
      if (q0 < p.Lq && q < p.Lq) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.Lq + q) * p.ldo + h * kP2D + hf * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(0, 0, 0, 0);
        if (p.lse2 && hf == 0) p.lse2[((long long)b * p.H + h) * p.Lq + q] = 0.f;
      }
    } else {
      for (int j = 0; j < n_g; ++j) {
        const int k0 = j * kP2TK;
        const int kh = k0 + hf * 64;
        const bool tr = (wg == 0 && lane == 0);
        if (tr) VC_TRACE2(g * 512 + j * 8 + 0, 1 + g * 20);
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        if (tr) VC_TRACE2(g * 512 + j * 8 + 1, 2 + g * 20);
        // ---- pass 1: row max of s2 over the tile (TMEM reads are cheap: ~900 B/clk/SM measured, the tile is 64 KB)
        const bool causal_tile = p.causal && (k0 + kP2TK - 1 > q0);
        // uniform over the group: the tile's bias window is one value and every key attends
        const bool fast = !causal_tile && (sFlagB[j + 1 - g] | sFlagB[j + 2 - g] | sFlagP[j]) == 0;
        const float cb = sBias[k0 + 255 - g * kP2TQ];
        float m_loc = -INFINITY;
        if (fast) {
          float v[32], w[32];
          tmem_ld32(tS, v);
          tmem_ld32(tS + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) m_loc = fmaxf(m_loc, fmaxf(v[i], w[i]));
          m_loc = fmaf(m_loc, p.scale_log2e, cb);   // scale > 0: max commutes with the affine map
        } else {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld32(tS + c * 32, v);
            const float4* pen4 = reinterpret_cast<const float4*>(sPen + kh + c * 32);
            const float* bk = brow + kh + c * 32;
            const int tq = q - kh - c * 32;  // column i is causally masked iff i > tq
            tmem_ld_wait();
            // s2 (+ causal) is written back to TMEM so that pass 2 only subtracts the max and exponentiates
            if (causal_tile) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 pe = pen4[i >> 2];
                const float pen[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float s2 = fminf(fmaf(v[i + e], p.scale_log2e, bk[i + e]), pen[e]);
                  s2 = (i + e > tq) ? fminf(s2, kP2Masked) : s2;
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
            tmem_st32(tS + c * 32, v);
          }
          tmem_st_wait();
        }
        // exchange the half-row maxima (slots double-buffered by tile parity)
        float* mx = mxg + (j & 1) * 256;
        mx[hf * 128 + r] = m_loc;
        if (tr) VC_TRACE2(g * 512 + j * 8 + 2, 3 + g * 20);
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        if (tr) VC_TRACE2(g * 512 + j * 8 + 3, 4 + g * 20);
        const float m_new = fmaxf(m_run, ceilf(fmaxf(m_loc, mx[(hf ^ 1) * 128 + r])));
        const float corr = fast_exp2(m_run - m_new);   // first tile: m_run = -inf -> 0
        if (j > 0) {
          mbar_wait(&pv_done[g], (j - 1) & 1);         // PV_g(j-1) complete: O_g is stable, sP_g may be overwritten
          tc_fence_after();
          if (__any_sync(0xffffffffu, m_new > m_run)) {   // rare after the first tiles
            float o[32];
            tmem_ld32(tO, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= corr;
            tmem_st32(tO, o);
            tmem_st_wait();
          }
        }
        if (tr) VC_TRACE2(g * 512 + j * 8 + 4, 5 + g * 20);
        // ---- pass 2: p = exp2(s2 - m_new) -> bf16 P (swizzled K-major A operand), row sum.  Fast tiles still hold the
        // raw accumulator: one FMA folds scale, bias and the max.
        const float e_mul = fast ? p.scale_log2e : 1.0f;
        const float e_add = fast ? cb - m_new : -m_new;
        float l_tile = 0.0f;
        uint8_t* prow = sPg + hf * 16384 + r * 128;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(tS + c * 32, v);
          tmem_ld_wait();
          if (c == 1) {
            // this warp has read the last of S_g(j): release the buffer so that the MMA warp can issue S_g(j+1)
            // underneath the rest of this tile's exponentials
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pv = fast_exp2(fmaf(v[i], e_mul, e_add));
            l_tile += pv;                  // the normaliser uses the un-dropped probabilities
            v[i] = pv;
          }
          // dropout: kept probabilities stay unscaled, 1/(1-p) is folded into the final normalisation
          if (p.drop_p16) drop_select<32>(v, drop_rk, p.drop_p16, (uint32_t)(kh + c * 32));
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const int ch = (c * 4 + gg) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + ch * 16) =
                make_uint4(pack_bf16x2(v[gg * 8 + 0], v[gg * 8 + 1]), pack_bf16x2(v[gg * 8 + 2], v[gg * 8 + 3]),
                           pack_bf16x2(v[gg * 8 + 4], v[gg * 8 + 5]), pack_bf16x2(v[gg * 8 + 6], v[gg * 8 + 7]));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        if (tr) VC_TRACE2(g * 512 + j * 8 + 5, 6 + g * 20);
        l_run = l_run * corr + l_tile;
        m_run = m_new;
      }
      // ---- epilogue: combine the half-row sums, normalise O and store
      float* sLg = sL + g * 256;
      sLg[hf * 128 + r] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      const float l_row = l_run + sLg[(hf ^ 1) * 128 + r];
      mbar_wait(&pv_done[g], (n_g - 1) & 1);
      tc_fence_after();
      float o[32];
      tmem_ld32(tO, o);
      tmem_ld_wait();
      if (q < p.Lq) {
        const float inv = (p.drop_p16 ? drop_scale(p.drop_p16) : 1.0f) / l_row;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.Lq + q) * p.ldo + h * kP2D + hf * 32);
#pragma unroll
        for (int gg = 0; gg < 4; ++gg)
          dst[gg] = make_uint4(pack_bf16x2(o[gg * 8 + 0] * inv, o[gg * 8 + 1] * inv),
                               pack_bf16x2(o[gg * 8 + 2] * inv, o[gg * 8 + 3] * inv),
                               pack_bf16x2(o[gg * 8 + 4] * inv, o[gg * 8 + 5] * inv),
                               pack_bf16x2(o[gg * 8 + 6] * inv, o[gg * 8 + 7] * inv));
        if (p.lse2 && hf == 0) p.lse2[((long long)b * p.H + h) * p.Lq + q] = m_run + log2f(l_row);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vc

using namespace vc;

namespace vc {
// Called by vc_attn_fwd (attn_fwd.cu) for training shapes (Lq > 128, no incremental-decoding fields).
int launch_attn_fwd_pair(const vc_attn_args* a, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  int s;
  if ((s = make_tmap_3d(&tmQ, a->q, a->ldq, a->Lq, a->B, a->ldq, (uint64_t)a->Lq * a->ldq, 64, kP2TQ)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmK, a->k, a->ldk, a->Lk, a->B, a->ldk, (uint64_t)a->Lk * a->ldk, 64, kP2TK)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmV, a->v, a->ldv, a->Lk, a->B, a->ldv, (uint64_t)a->Lk * a->ldv, 64, kP2TK)) != VC_OK) return s;
  AttnFwd2Params p;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.q_col = a->q_col; p.k_col = a->k_col; p.v_col = a->v_col;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  p.lse2 = a->lse2; p.bias_rel = a->bias_rel; p.kmask = a->kmask; p.causal = a->causal;
  p.q_like_k = a->q_like_k;
  p.scale_log2e = a->scale * kP2Log2e;
  p.drop_seed = a->drop_seed; p.drop_p16 = a->drop_p16; p.drop_salt = drop_salt_ptr();
  p.trace = debug_trace_ptr();
  VC_CHECK(a->Lk <= kP2MaxLk, "vc_attn_fwd: Lk=%d exceeds the %d keys the bias/mask staging supports", a->Lk, kP2MaxLk);
  const int lk_pad = ((a->Lk + kP2TK - 1) / kP2TK) * kP2TK;
  static PerDeviceOnce attr;
  if (attr.need()) {
    VC_CUDA(cudaFuncSetAttribute(attn_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_fwd2_smem(kP2MaxLk)));
  }
  dim3 grid((a->Lq + 2 * kP2TQ - 1) / (2 * kP2TQ), a->H, a->B);
  VC_CUDA(launch_kernel(attn_fwd_pair_kernel, grid, dim3(576), (size_t)attn_fwd2_smem(lk_pad), st, tmQ, tmK, tmV, p));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
}  // namespace vc
