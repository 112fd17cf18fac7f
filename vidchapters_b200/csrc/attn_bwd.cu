// Fused attention backward on tcgen05 (sm_100a), head_dim 64: dQ, dK, dV and the relative-position-bias gradient.
//
// Replaces autograd through modeling_t5.py:539-580 / vit.py:47-51.  Probabilities are recomputed from the saved
// log-sum-exp (flash-attention style); nothing of size Lq x Lk touches HBM.
//
// One CTA = one (batch, head, 128-key tile), looping over the 128-query tiles that see it.  1 CTA / SM
// (215 KB smem, 448 TMEM columns).  64 + 32 NCW threads (NCW = 16 compute warps by default):
//   warp 0 lane 0 : TMA: K,V tile once; (Q_i, dO_i) through a 2-stage ring
//   warp 1        : TMEM alloc; lane 0 issues per query tile
//                     S  = Q_i.K^T        dP = dO_i.V^T                       (128x128x64 each, fresh)
//                     dV += P^T.dO_i      dK += dS^T.Q_i     dQ_i = dS.K      (accumulate / accumulate / fresh)
//   warps 2..     : compute warps (NCW/4 per TMEM lane quarter, one or two 32-column chunks each): p = exp2(s2 - lse2),
//                   ds = p*(dP - delta); write bf16 P and scale*dS tiles (128B-swizzled; each tile serves as MN-major
//                   A for dV/dK and dS also as K-major A for dQ).  Tiles whose bias window is one value and whose keys
//                   all attend skip every per-element lookup; key tiles made only of padding exit at once.
//                   d(bias): one sum per tile when the tile lies in one bucket, else per-diagonal sums by warp
//                   shuffles.  dQ_i goes TMEM -> swizzled smem -> TMA reduce-add into the fp32 dQ accumulator
//                   (warps 2..9); finally dK, dV are stored (bf16).
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

constexpr int kBT = 128, kBD = 64;
constexpr float kBLog2e = 1.4426950408889634f;
constexpr float kBMasked = -3.0e38f;

struct AttnBwdParams {
  int B, H, Lq, Lk;
  int q_col, k_col, v_col, do_col;
  const float* lse2;      // [B,H,Lq]
  const float* delta;     // [B,H,Lq] rowsum(dO*O)
  const float* bias_rel;  // [H][Lq+Lk-1] or null
  const int* bucket_lut;  // [Lq+Lk-1] bucket id per relative position (uniform-tile test), or null
  const uint8_t* kmask;   // [B][Lk] or null
  int causal;
  float scale, scale_log2e, inv_scale;
  float* dq_acc;          // fp32 [B*Lq][ld_dq], head h at cols 64h   (atomicAdd)
  long long ld_dq;
  __nv_bfloat16* dk; long long ld_dk; int dk_col;  // bf16 [B*Lk][ld], head h at cols dk_col + 64h
  __nv_bfloat16* dv; long long ld_dv; int dv_col;
  float* dbias_rel;       // fp32 [H][Lq+Lk-1] (atomicAdd) or null
  uint32_t drop_seed, drop_p16;
  const uint32_t* drop_salt;
  int q_like_k;           // self-attention over a padded sequence: query tiles past the last attended key are skipped
  long long* trace;       // debug (vc_debug_set_trace): clock64 timeline of CTA (0,0,0), warp 2 lane 0 and the MMA thread
};

#define VC_TRACE(slot, ev)                                                                                   \
  do {                                                                                                       \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (slot) < 2048)                    \
      p.trace[(slot)] = ((long long)(ev) << 48) | (clock64() & 0xFFFFFFFFFFFFLL);                             \
  } while (0)

constexpr int kBwdSmemTiles = 16384 * 2 /*K,V*/ + 2 * 16384 /*Q ring*/ + 2 * 16384 /*dO ring*/ + 32768 /*P*/ + 32768 /*dS*/;
constexpr int kBwdRelMax = 2304;  // floats per relative-position window: ceil(Lq/128)*128 + 128 <= 2304 (Lq <= 2176)
constexpr int kBwdTsum = 16 * 24 * 4;   // per compute warp x query tile: d(bias) total of a one-bucket tile
constexpr int kBwdStat = 2 * 512 * 8;   // per compute thread x 2 stages: (lse2, delta) of its query row
constexpr int kAttnBwdSmem = kBwdSmemTiles + 32768 + 2 * kBwdRelMax * 4 + 512 + 512 + kBwdTsum + kBwdStat;  // + dQ staging + d(bias)/bias windows + key ceilings

template <int NCW>   // compute warps: 8 (two 32-column chunks of the tile per thread) or 16 (one chunk per thread)
__global__ void __launch_bounds__(64 + NCW * 32, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + 16384;
  uint8_t* sQ = sV + 16384;        // [2]
  uint8_t* sDO = sQ + 2 * 16384;   // [2]
  uint8_t* sP = sDO + 2 * 16384;   // 32 KB
  uint8_t* sDS = sP + 32768;       // 32 KB
  uint8_t* sDQ = sDS + 32768;      // 32 KB: dQ_i staged as two 128-row x 32-col fp32 boxes (SWIZZLE_128B) for TMA reduce-add
  float* s_rel = reinterpret_cast<float*>(sDQ + 32768);   // d(bias) window of this key tile: slot (k - k0) + (Lq - 1 - q)
  float* s_bias = s_rel + kBwdRelMax;       // bias(k - q) * log2e, same slot indexing
  float* s_pen = s_bias + kBwdRelMax;   // per-key ceiling of this tile: +inf attend | kBMasked | -inf out of range
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_pen + 128);
  uint64_t* kv_full = bars + 0;
  uint64_t* qd_full = bars + 1;   // [2]
  uint64_t* qd_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_read = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  // per query tile: the bias over the whole (query tile x this key tile) window is not one value / not one bucket
  int* s_nonuni_v = reinterpret_cast<int*>(bars + 10);   // [24]
  int* s_nonuni_b = s_nonuni_v + 24;                      // [24]
  int* s_anypen = s_nonuni_b + 24;                        // [1] some key of this tile is masked / out of range
  int* s_tile_att = s_anypen + 1;                         // [1] some key of this tile attends
  int* s_row_att = s_anypen + 2;                          // [1] some key of this batch row attends (causal: key 0 does)
  int* s_last_key = s_anypen + 3;                         // [1] last attended key of this batch row (q_like_k), -1: none
  float* s_tsum = reinterpret_cast<float*>(bars + 64);    // [16 warps][24 query tiles]
  float2* s_stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(s_tsum) + kBwdTsum);   // [2][512]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * kBT;

  if (threadIdx.x == 0) VC_TRACE(1900, 200);
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(pds_full, NCW);   // one arrival per compute warp
    mbar_init(dq_full, 1);
    mbar_init(dq_read, NCW == 16 ? 16 : 8);   // one arrival per dQ-staging warp
    fence_barrier_init();
  }
  const int nqt_all = (p.Lq + kBT - 1) / kBT;
  const int qt0 = p.causal ? min(kt, nqt_all) : 0;  // causal: query tiles before the key tile see nothing of it
  if (threadIdx.x < 51) s_nonuni_v[threadIdx.x] = 0;   // (all flag arrays)
  if (threadIdx.x == 51) *s_last_key = -1;
  if (warp == 1) tmem_alloc(tmem_slot, 512);   // before the dependency wait: touches no global memory
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is read from here on
  pdl_trigger();
  if (threadIdx.x == 0) VC_TRACE(1901, 201);
  // The key/value tile and the first query tile are requested BEFORE the mask / bias set-up below, so their L2 latency
  // runs underneath it (with only two query tiles per CTA — cross-attention — the set-up was ~half of the CTA's life).
  // Every CTA has at least one query tile unless the causal mask hides them all (qt0 == nqt_all).
  const bool early_q = qt0 < nqt_all;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(kv_full, 32768);
    tma_load_3d(sK, &tmK, kv_full, p.k_col + h * kBD, k0, b);
    tma_load_3d(sV, &tmV, kv_full, p.v_col + h * kBD, k0, b);
    if (early_q) {
      mbar_arrive_expect_tx(&qd_full[0], 32768);
      tma_load_3d(sQ, &tmQ, &qd_full[0], p.q_col + h * kBD, qt0 * kBT, b);
      tma_load_3d(sDO, &tmDO, &qd_full[0], p.do_col + h * kBD, qt0 * kBT, b);
    }
  }
  const int n_rel = p.Lq + kBT;  // relative positions touched by this CTA: (k - q + Lq - 1) - rel_base in [0, Lq+127)
  const int rel_base = k0;       // slot = (k - k0) + (Lq - 1 - q)
  // Set-up by the compute warps only (threads 64..): warp 0's elected thread is busy issuing the loads above and the
  // CTA-wide barrier below waits for the slowest thread.  The mask bytes are requested before anything is stored so the
  // global round trips overlap.
  const int su = (int)threadIdx.x - 64, sn = (int)blockDim.x - 64;
  if (su >= 0) {
    const unsigned char* km = p.kmask ? p.kmask + (long long)b * p.Lk : nullptr;
    unsigned char my_key = 1;
    if (su < 128 && km && k0 + su < p.Lk) my_key = km[k0 + su];
    int last = -1;
    if (km && !p.causal) {
      for (int i = su; i < p.Lk; i += sn)
        if (km[i] != 0) last = i;
    }
    const bool row0_att = (km && p.causal && su == 0) ? km[0] != 0 : false;
    if (p.dbias_rel) {
      for (int i = su; i < kBwdRelMax; i += sn) s_rel[i] = 0.f;
      for (int i = su; i < 16 * 24; i += sn) s_tsum[i] = 0.f;
    }
    // always filled + padded so the per-element loop below is branch-free (see attn_fwd.cu)
    const float* brow_g = p.bias_rel ? p.bias_rel + (long long)h * (p.Lq + p.Lk - 1) : nullptr;
    const int n_pad = ((p.Lq + kBT - 1) / kBT) * kBT + kBT;  // covers slot (k-k0) + (Lq-1-q) for every q of every tile
    for (int i = su; i < n_pad; i += sn)
      s_bias[i] = (brow_g && rel_base + i < p.Lq + p.Lk - 1) ? __ldg(brow_g + rel_base + i) * kBLog2e : 0.f;
    if (su < 128) {
      const int k = k0 + su;
      const float pen = (k >= p.Lk) ? -INFINITY : (my_key == 0 ? kBMasked : INFINITY);
      s_pen[su] = pen;
      if (pen != INFINITY) *s_anypen = 1;
      else *s_tile_att = 1;
    }
    if (!km) {
      if (su == 0) *s_row_att = 1;
    } else if (p.causal) {
      if (row0_att) *s_row_att = 1;
    } else {
      last = __reduce_max_sync(0xffffffffu, last);
      if (lane == 0 && last >= 0) { *s_row_att = 1; atomicMax(s_last_key, last); }
    }
  }
  // Uniform-tile flags.  Query tile t touches window slots [max(w0,0), w0+254], w0 = Lq-128-128t.  With T5's buckets every
  // tile further than 128 positions from the diagonal sees ONE bias value (one bucket): s2 = acc*scale + c without any
  // per-element lookup, and d(bias) of the tile is one sum instead of 255 diagonal sums.  (Read from global memory, not
  // from s_bias, so that the set-up needs a single CTA-wide barrier; without a bias every tile is uniform.)
  if (p.bias_rel) {
    const float* brow_g = p.bias_rel + (long long)h * (p.Lq + p.Lk - 1);
    const int lim = p.Lq + p.Lk - 1;
    for (int idx = su; idx >= 0 && idx < nqt_all * 256; idx += sn) {
      const int t = idx >> 8, w0 = p.Lq - kBT - t * kBT;
      const int i = max(w0, 0) + 1 + (idx & 255);
      if (i <= w0 + 254) {
        const int gi = rel_base + i;
        const float b1 = gi < lim ? __ldg(brow_g + gi) : 0.f, b0 = gi - 1 < lim ? __ldg(brow_g + gi - 1) : 0.f;
        if (b1 != b0) s_nonuni_v[t] = 1;
        if (p.dbias_rel) {
          if (!p.bucket_lut || gi >= lim || __ldg(p.bucket_lut + gi) != __ldg(p.bucket_lut + gi - 1)) s_nonuni_b[t] = 1;
        }
      }
    }
  }
  __syncthreads();
  if (*s_tile_att == 0 && *s_row_att != 0) {
    // Every key of this tile is masked (padding) while each query row attends to some other key: the tile's
    // probabilities are exp2(finfo.min - lse) == 0 exactly, so dK = dV = 0 and it adds nothing to dQ / d(bias).
    const int half = (threadIdx.x >> 7) & 1, r = threadIdx.x & 127, kk = k0 + r;
    if (threadIdx.x < 256 && kk < p.Lk) {
      uint4* d1 = reinterpret_cast<uint4*>(p.dv + ((long long)b * p.Lk + kk) * p.ld_dv + p.dv_col + h * kBD + half * 32);
      uint4* d2 = reinterpret_cast<uint4*>(p.dk + ((long long)b * p.Lk + kk) * p.ld_dk + p.dk_col + h * kBD + half * 32);
      for (int g = 0; g < 4; ++g) { d1[g] = make_uint4(0, 0, 0, 0); d2[g] = make_uint4(0, 0, 0, 0); }
    }
    if (threadIdx.x == 0) {   // the early loads must have landed before this CTA's shared memory is given away
      mbar_wait(kv_full, 0);
      if (early_q) mbar_wait(&qd_full[0], 0);
    }
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc(*tmem_slot, 512);
    }
    return;
  }
  if (threadIdx.x == 0) VC_TRACE(1902, 202);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;

  // q_like_k: query rows past the sequence's last token are padding whose upstream gradient is exactly zero (nothing
  // downstream can observe them), so their tiles add nothing to dK / dV / d(bias) and their dQ stays zero: skip them
  const int nqt_end = (p.q_like_k && *s_last_key >= 0) ? min(nqt_all, *s_last_key / kBT + 1) : nqt_all;
  const int nqt = max(nqt_end - qt0, 0);

  if (warp == 0 && lane == 0) {
    for (int i = 1; i < nqt; ++i) {   // (tile 0 and K/V were requested at the top)
      const int st = i & 1;
      mbar_wait(&qd_empty[st], ((i >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx(&qd_full[st], 32768);
      tma_load_3d(sQ + st * 16384, &tmQ, &qd_full[st], p.q_col + h * kBD, (qt0 + i) * kBT, b);
      tma_load_3d(sDO + st * 16384, &tmDO, &qd_full[st], p.do_col + h * kBD, (qt0 + i) * kBT, b);
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);   // S, dP : A K-major, B K-major, N=128
    constexpr uint32_t id_t = make_idesc_bf16(128, 64, 1, 1);    // dV, dK: A MN-major (P^T / dS^T), B MN-major, N=64
    constexpr uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);    // dQ    : A K-major (dS), B MN-major (K), N=64
    mbar_wait(kv_full, 0);
    VC_TRACE(1903, 203);
    const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK), 0, 1024);
    const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV), 0, 1024);
    // Software pipeline: S/dP of tile i+1 are issued BEFORE dV/dK/dQ of tile i (the score buffers are free as soon as the
    // compute warps have written P/dS(i), i.e. at pds_full(i)), so the compute warps run the register math of tile i+1
    // underneath the 768 cycles of dV/dK/dQ(i) instead of waiting for them.
    auto issue_scores = [&](int i) {
      const int st = i & 1;
      mbar_wait(&qd_full[st], (i >> 1) & 1);
      tc_fence_after();
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + st * 16384), 0, 1024);
      const uint64_t dodesc = make_smem_desc_sw128(smem_u32(sDO + st * 16384), 0, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tS, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), id_s, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tDP, dodesc + (uint64_t)(k * 2), vdesc + (uint64_t)(k * 2), id_s, k > 0);
      tc_commit(s_full);
    };
    if (nqt > 0) issue_scores(0);
    for (int i = 0; i < nqt; ++i) {
      const int st = i & 1;
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + st * 16384), 0, 1024);
      const uint64_t dodesc = make_smem_desc_sw128(smem_u32(sDO + st * 16384), 0, 1024);
      VC_TRACE(1024 + i * 8 + 0, 100);
      mbar_wait(pds_full, i & 1);          // P/dS(i) in shared memory; S(i), dP(i) consumed
      tc_fence_after();
      VC_TRACE(1024 + i * 8 + 1, 101);
      if (i + 1 < nqt) issue_scores(i + 1);
      VC_TRACE(1024 + i * 8 + 2, 102);
      if (i > 0) mbar_wait(dq_read, (i - 1) & 1);   // dQ(i-1) has left tensor memory
      tc_fence_after();
      VC_TRACE(1024 + i * 8 + 3, 103);
      // MN-major A over the [q rows][kv cols] tiles: 2 atoms of 64 kv (LBO = 16384), 8-row groups 1024 B, +2048 B per 16 q.
      const uint64_t pT = make_smem_desc_sw128(smem_u32(sP), 16384, 1024);
      const uint64_t dsT = make_smem_desc_sw128(smem_u32(sDS), 16384, 1024);
#pragma unroll
      for (int k = 0; k < 8; ++k) tc_mma_bf16(tDV, pT + (uint64_t)(k * 128), dodesc + (uint64_t)(k * 128), id_t, (i > 0 || k > 0));
#pragma unroll
      for (int k = 0; k < 8; ++k) tc_mma_bf16(tDK, dsT + (uint64_t)(k * 128), qdesc + (uint64_t)(k * 128), id_t, (i > 0 || k > 0));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t dsK = make_smem_desc_sw128(smem_u32(sDS + (k >> 2) * 16384) + (k & 3) * 32, 0, 1024);
        tc_mma_bf16(tDQ, dsK, kdesc + (uint64_t)(k * 128), id_q, k > 0);
      }
      tc_commit(&qd_empty[st]);
      tc_commit(dq_full);
      VC_TRACE(1024 + i * 8 + 4, 104);
    }
  } else if (warp >= 2) {
    // ===================== compute: 8 warps = 2 per TMEM lane quarter, each taking 2 of the 4 column chunks ==========
    const int quarter = warp & 3;
    constexpr int CPT = 16 / NCW;            // 32-column chunks per thread (4 per row, NCW/4 warps per row quarter)
    const int part = (warp - 2) >> 2;        // which chunk group of the 128 keys
    const int half = part & 1;               // which 32 of the 64 head columns (dQ staging, dV/dK epilogue: warps 2..9)
    const bool io_warp = part < 2;
    const int r = quarter * 32 + lane;
    const int ct = (warp - 2) * 32 + lane;   // 0..255: index among the compute threads
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float drop_sc = p.drop_p16 ? drop_scale(p.drop_p16) : 1.0f;
    static_assert(NCW == 16 || NCW == 8, "compute warps");
    // dQ_i: TMEM -> swizzled smem staging -> ONE TMA reduce-add per 32-column half (whole 128-byte lines into the fp32
    // dQ accumulator) instead of 2048 scattered 16-byte atomics per tile.  Called by warps 2..9 for tile t once
    // dq_full(t) has been observed.
    auto stage_dq = [&](int t) {
      const int q0s = (qt0 + t) * kBT;
      if constexpr (NCW == 16) {
        // Warp-local: every compute warp moves ITS 32 rows x 16 columns of dQ: TMEM -> 2 KB private staging block
        // (SWIZZLE_64B) -> one TMA reduce-add (box 16 x 32 fp32) into the fp32 dQ accumulator.  No CTA-wide barrier: with
        // two bar.syncs over a shared 128-row box the staging cost ~1500 cycles per tile, all of it waiting for the
        // slowest warp (r02 timeline, profiles/r02_attn_bwd_timeline.txt).
        float v[16];
        tmem_ld16(tDQ + lane_off + part * 16, v);
        if (lane == 0) bulk_wait_read0();           // this warp's previous reduce has finished reading its block
        __syncwarp();
        tmem_ld_wait();
        if (warp == 2 && lane == 0) VC_TRACE(512 + t * 4 + 0, 8);
        uint8_t* blk = sDQ + (warp - 2) * 2048;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<float4*>(blk + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
              make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (warp == 2 && lane == 0) VC_TRACE(512 + t * 4 + 1, 9);
        if (lane == 0) {   // rows past Lq carry zeros (p = 0 there); rows past the tensor are clipped by TMA
          tma_reduce_add_2d(&tmDQ, blk, h * kBD + part * 16, b * p.Lq + q0s + quarter * 32);
          bulk_commit();
        }
        if (warp == 2 && lane == 0) VC_TRACE(512 + t * 4 + 2, 10);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_read);
        return;
      } else {
        float v[32];
        tmem_ld32(tDQ + lane_off + half * 32, v);
        if (ct == 0) bulk_wait_read0();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        tmem_ld_wait();
        uint8_t* drow_q = sDQ + half * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(drow_q + ((g ^ (r & 7)) << 4)) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      if (ct == 0) {   // rows past Lq carry zeros (p = 0 there); rows past the tensor are clipped by TMA
        tma_reduce_add_2d(&tmDQ, sDQ, h * kBD, b * p.Lq + q0s);
        tma_reduce_add_2d(&tmDQ, sDQ + 16384, h * kBD + 32, b * p.Lq + q0s);
        bulk_commit();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_read);
    };
    // Per-row statistics (lse2, delta) of the next query tile are fetched one tile ahead — their L2 latency was 14 % of
    // the kernel's stall samples when loaded at the top of the tile (profiles/r02_attn_source_counters.md) — and by
    // cp.async into the thread's own shared-memory slot: held in registers across a tile, ptxas (96 registers at 576
    // threads) spilled them right behind the load, which waits for the load after all (~800 cycles per tile).
    auto fetch_stats = [&](int tile_q0, int slot) {
      const int qn = tile_q0 + r;
      float* dst = reinterpret_cast<float*>(&s_stat[slot * 512 + ct]);
      if (qn < p.Lq) {
        const long long si = ((long long)b * p.H + h) * p.Lq + qn;
        cp_async4(dst, p.lse2 + si);
        cp_async4(dst + 1, p.delta + si);
      } else {
        dst[0] = INFINITY;   // rows past Lq: p = exp2(-inf) = 0
        dst[1] = 0.f;
      }
      cp_async_commit();
    };
    if (nqt > 0) fetch_stats(qt0 * kBT, 0);
    for (int i = 0; i < nqt; ++i) {
      const int q0 = (qt0 + i) * kBT;
      const int q = q0 + r;
      const bool q_ok = q < p.Lq;
      cp_async_wait_all();
      const float2 stat = s_stat[(i & 1) * 512 + ct];
      const float lse2 = stat.x;
      const float delta = stat.y;
      if (i + 1 < nqt) fetch_stats(q0 + kBT, (i + 1) & 1);
      // indexed by k - k0; rows past Lq (zero-filled Q/dO, p forced to 0) clamp to slot 0 to stay inside the window
      const float* brow = s_bias + (q_ok ? (p.Lq - 1 - q) : 0);
      const bool causal_tile = p.causal && (k0 + kBT - 1 > q0);
      // uniform over the CTA (flags per query tile): constant bias, every key attends, no causal edge
      const bool fast = !causal_tile && (s_nonuni_v[qt0 + i] | *s_anypen) == 0;
      const bool tile_sum = p.dbias_rel && !causal_tile && (s_nonuni_b[qt0 + i] | *s_anypen) == 0 && fast;
      const float c_fast = s_bias[max(p.Lq - kBT - q0, 0) + 1] - lse2;
      const float nds = -delta * p.scale;          // -delta * scale
      const float sc_s = drop_sc * p.scale;        // dropout 1/(1-p) * scale
      float ds_sum = 0.f;
      if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 0, 1);
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 1, 2);
#pragma unroll 1
      for (int cc = 0; cc < CPT; ++cc) {
        const int c = part * CPT + cc;
        float sv[32], dp[32];
        tmem_ld32(tS + lane_off + c * 32, sv);
        tmem_ld32(tDP + lane_off + c * 32, dp);
        const float4* pen4 = reinterpret_cast<const float4*>(s_pen + c * 32);
        const float* bk = brow + c * 32;
        const int tq = q - k0 - c * 32;  // column j is causally masked iff j > tq
        tmem_ld_wait();
        if (fast) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sv[j] = fast_exp2(fmaf(sv[j], p.scale_log2e, c_fast));
        } else if (causal_tile) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 pe = pen4[j >> 2];
            const float pen[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float s2 = fminf(fmaf(sv[j + e], p.scale_log2e, bk[j + e]), pen[e]);
              s2 = (j + e > tq) ? fminf(s2, kBMasked) : s2;
              sv[j + e] = fast_exp2(s2 - lse2);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 pe = pen4[j >> 2];
            const float pen[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              sv[j + e] = fast_exp2(fminf(fmaf(sv[j + e], p.scale_log2e, bk[j + e]), pen[e]) - lse2);
          }
        }
        // ds (already multiplied by `scale`, the factor dK/dQ need) = p * scale * (dP - delta); with dropout
        // O = (mask.P).V * sc: dV accumulates the masked, UNSCALED P (sc is applied once when dV is stored) and dP flows
        // back through the same mask: ds = p * scale * (keep ? dP*sc - delta : -delta).
        if (p.drop_p16) {
          const uint32_t rk = drop_row_key(drop_salted(p.drop_seed, p.drop_salt), ((unsigned long long)b * p.H + h) * p.Lq + q);
          const uint32_t hsh = drop_block_hash(rk, (uint32_t)(k0 + c * 32)), thr = drop_threshold(p.drop_p16);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const bool keep = drop_keep_h(hsh, j, thr);
            dp[j] = sv[j] * (keep ? fmaf(dp[j], sc_s, nds) : nds);
            sv[j] = keep ? sv[j] : 0.0f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dp[j] = sv[j] * fmaf(dp[j], p.scale, nds);
        }
        if (tile_sum) {
#pragma unroll
          for (int j = 0; j < 32; ++j) ds_sum += dp[j];
        } else if (p.dbias_rel) {
          // d(bias)[k - q] of this warp's 32x32 block: element (row l', col j) lies on diagonal j - l'.  Lane L
          // collects the diagonals congruent to L (mod 32): for each j one shuffle from lane (j - L) mod 32 brings it
          // the element of diagonal L (j >= L) or L - 32 (j < L).  No shared-memory traffic, no load imbalance.
          float a_pos = 0.f, a_neg = 0.f;
          const int neg_l = 32 - lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float w = __shfl_sync(0xffffffffu, dp[j], neg_l + j);   // source lane taken modulo 32
            if (j >= lane) a_pos += w; else a_neg += w;
          }
          const int slot = c * 32 - quarter * 32 + (p.Lq - 1 - q0) + lane;   // (k - k0) + (Lq - 1 - q) of diagonal +lane
          if (slot >= 0 && a_pos != 0.f) atomicAdd(&s_rel[slot], a_pos * p.inv_scale);   // d(bias) wants the unscaled ds
          if (slot >= 32 && a_neg != 0.f) atomicAdd(&s_rel[slot - 32], a_neg * p.inv_scale);
        }
        // pack first: while this tile's math ran, the MMAs of the PREVIOUS tile (dV, dK, dQ) may still have been
        // reading the P / dS tiles; only the stores below have to wait for them
        uint32_t pk[16], dk_[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { pk[j] = pack_bf16x2(sv[2 * j], sv[2 * j + 1]); dk_[j] = pack_bf16x2(dp[2 * j], dp[2 * j + 1]); }
        if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 2, 3);
        if (cc == 0 && i > 0) {
          mbar_wait(dq_full, (i - 1) & 1);
          tc_fence_after();
        }
        if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 3, 4);
        uint8_t* prow = sP + (c >> 1) * 16384 + r * 128;
        uint8_t* drow = sDS + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = (((c & 1) * 4 + g) ^ (r & 7)) * 16;
          *reinterpret_cast<uint4*>(prow + ch) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
          *reinterpret_cast<uint4*>(drow + ch) = make_uint4(dk_[g * 4], dk_[g * 4 + 1], dk_[g * 4 + 2], dk_[g * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 4, 5);
      if (tile_sum) {
        // one bucket for the whole tile: deposit the tile's total at one of its relative positions (k = k0, q = q0)
        // (a plain store into this warp's slot: 16 warps hitting one shared-memory word with float atomics — CAS
        //  loops — cost ~500 cycles per tile in the r02 timeline; the slots are folded into s_rel at the end)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ds_sum += __shfl_xor_sync(0xffffffffu, ds_sum, o);
        if (lane == 0) s_tsum[(warp - 2) * 24 + (qt0 + i)] = ds_sum * p.inv_scale;
      }
      // dQ of the PREVIOUS tile (complete: dq_full(i-1) was observed above) leaves while the tensor pipe computes the
      // scores of tile i+1
      if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 6, 7);
      if ((io_warp || NCW == 16) && i > 0) stage_dq(i - 1);
      if (warp == 2 && lane == 0) VC_TRACE(i * 8 + 5, 6);
    }
    if ((io_warp || NCW == 16) && nqt > 0) {
      mbar_wait(dq_full, (nqt - 1) & 1);     // also covers the final dV / dK MMAs
      tc_fence_after();
      stage_dq(nqt - 1);
    }
    if (warp == 2 && lane == 0) VC_TRACE(1904, 204);
    // ---- dV, dK: rows = keys of this tile.  (The last dq_full wait also covers the final dV/dK MMAs.)  16 compute warps:
    // every warp stores its 32 rows x 16 columns of both; 8 warps: warps 2..9 store 32 columns each.
    const int kk = k0 + r;
    if (NCW == 16 && nqt > 0) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
        const long long ld = which == 0 ? p.ld_dv : p.ld_dk;
        const int col = which == 0 ? p.dv_col : p.dk_col;
        float v[16];
        tmem_ld16((which == 0 ? tDV : tDK) + lane_off + part * 16, v);
        tmem_ld_wait();
        if (which == 0 && p.drop_p16) {   // dropout scale of the probabilities, applied once per output instead of per element
#pragma unroll
          for (int g = 0; g < 16; ++g) v[g] *= drop_sc;
        }
        if (kk < p.Lk) {
          uint4* dst = reinterpret_cast<uint4*>(base + ((long long)b * p.Lk + kk) * ld + col + h * kBD + part * 16);
#pragma unroll
          for (int g = 0; g < 2; ++g)
            dst[g] = make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
        }
      }
    } else if (!io_warp) {
    } else if (nqt > 0) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
        const long long ld = which == 0 ? p.ld_dv : p.ld_dk;
        const int col = which == 0 ? p.dv_col : p.dk_col;
        float v[32];
        tmem_ld32((which == 0 ? tDV : tDK) + lane_off + half * 32, v);
        tmem_ld_wait();
        if (which == 0 && p.drop_p16) {   // dropout scale of the probabilities, applied once per output instead of per element
#pragma unroll
          for (int g = 0; g < 32; ++g) v[g] *= drop_sc;
        }
        if (kk < p.Lk) {
          uint4* dst = reinterpret_cast<uint4*>(base + ((long long)b * p.Lk + kk) * ld + col + h * kBD + half * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[g] = make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
        }
      }
    } else if (kk < p.Lk) {  // no query tile sees this key tile: zero gradients
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
        const long long ld = which == 0 ? p.ld_dv : p.ld_dk;
        const int col = which == 0 ? p.dv_col : p.dk_col;
        uint4* dst = reinterpret_cast<uint4*>(base + ((long long)b * p.Lk + kk) * ld + col + h * kBD + half * 32);
        for (int g = 0; g < 4; ++g) dst[g] = make_uint4(0, 0, 0, 0);
      }
    }
    // the dQ staging blocks only have to be READ by the TMA unit before the CTA exits (the adds complete with the grid);
    // waited for here, after the dV/dK stores, so that the two overlap
    if (NCW == 16 ? lane == 0 : ct == 0) bulk_wait_read0();
  }

  if (warp == 2 && lane == 0) VC_TRACE(1905, 205);
  tc_fence_before();
  __syncthreads();
  if (p.dbias_rel && threadIdx.x < nqt_all) {   // one-bucket tiles: total of the 16 warps -> the tile's slot (k = k0, q = q0)
    float t = 0.f;
    for (int w = 0; w < 16; ++w) t += s_tsum[w * 24 + threadIdx.x];
    if (t != 0.f) s_rel[p.Lq - 1 - threadIdx.x * kBT] += t;
  }
  __syncthreads();
  if (p.dbias_rel) {
    float* dst = p.dbias_rel + (long long)h * (p.Lq + p.Lk - 1) + rel_base;
    const int lim = p.Lq + p.Lk - 1 - rel_base;
    for (int i = threadIdx.x; i < n_rel && i < lim; i += blockDim.x) {
      const float g = s_rel[i];
      if (g != 0.f) atomicAdd(dst + i, g);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) VC_TRACE(1906, 206);
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (the "D" term of the softmax backward)
// One warp per row, rows grid-strided; a lane owns 8 consecutive elements of each 256-wide slab (4 heads) and issues the
// loads of all slabs of the row (H <= 16: up to 8 x 16 B per lane) before the first use.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout, long long lddo,
                  int do_col, float* __restrict__ delta, float* __restrict__ dq_acc, long long ld_dq, int B, int H, int Lq) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * Lq;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  const int nslab = (H + 3) >> 2;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const int b = (int)(row / Lq), q = (int)(row % Lq);
    uint4 a[4], g[4];
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      const int col = sl * 256 + lane * 8;
      const bool ok = sl < nslab && col < H * 64;
      a[sl] = ok ? *reinterpret_cast<const uint4*>(o + row * ldo + col) : make_uint4(0, 0, 0, 0);
      g[sl] = ok ? *reinterpret_cast<const uint4*>(dout + row * lddo + do_col + col) : make_uint4(0, 0, 0, 0);
    }
    // also clears this row of the fp32 dQ accumulator the main kernel adds into (saves a separate fill launch)
    for (int c = lane * 4; c < H * 64; c += 128) *reinterpret_cast<float4*>(dq_acc + row * ld_dq + c) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      if (sl < nslab) {
        const uint32_t aw[4] = {a[sl].x, a[sl].y, a[sl].z, a[sl].w}, gw[4] = {g[sl].x, g[sl].y, g[sl].z, g[sl].w};
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) s += bf16_lo(aw[e]) * bf16_lo(gw[e]) + bf16_hi(aw[e]) * bf16_hi(gw[e]);
        // reduce over the 8 lanes that share a head
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const int hh = sl * 4 + (lane >> 3);
        if ((lane & 7) == 0 && hh < H) delta[((long long)b * H + hh) * Lq + q] = s;
      }
    }
  }
}

}  // namespace vc

using namespace vc;

extern "C" int vc_attn_bwd(const vc_attn_bwd_args* a, void* stream) {
  VC_CHECK(a != nullptr, "vc_attn_bwd: null args");
  const vc_attn_args* f = &a->fwd;
  VC_CHECK(f->B > 0 && f->H > 0 && f->Lq > 0 && f->Lk > 0 && f->head_dim == 64, "vc_attn_bwd: bad dims");
  VC_CHECK(f->q_offset == 0 && f->q_offset_dev == nullptr && f->kv_batch_rows == 0 && f->bias_len == 0,
           "vc_attn_bwd: the incremental-decoding fields are forward-only");
  VC_CHECK(((f->Lq + kBT - 1) / kBT) * kBT + kBT <= kBwdRelMax, "vc_attn_bwd: Lq=%d too long for the d(bias) scratch", f->Lq);
  VC_CHECK(f->lse2 && a->delta && a->dq_acc && a->dk && a->dv && a->dout, "vc_attn_bwd: null buffers");
  VC_CHECK(a->ld_dq % 4 == 0 && a->ld_dk % 8 == 0 && a->ld_dv % 8 == 0 && a->ld_do % 8 == 0, "vc_attn_bwd: strides");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // delta = rowsum(dO * O)
  {
    const long long rows = (long long)f->B * f->Lq;
    VC_CHECK(f->H <= 16, "vc_attn_bwd: H=%d > 16", f->H);
    const long long want = (rows + 7) / 8, cap = (long long)num_sms() * 8;
    VC_CUDA(launch_kernel(attn_delta_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), 0, st, (const __nv_bfloat16*)f->out,
                          f->ldo, (const __nv_bfloat16*)a->dout, a->ld_do, a->do_col, a->delta, a->dq_acc, a->ld_dq, f->B, f->H,
                          f->Lq));
    VC_CUDA(cudaGetLastError());
  }
  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ;
  int s;
  VC_CHECK(((uintptr_t)a->dq_acc & 15) == 0, "vc_attn_bwd: dq_acc must be 16-byte aligned");
  static const int ncw = [] { const char* e = getenv("VIDCHAP_ATTN_BWD_WARPS"); return e && atoi(e) == 8 ? 8 : 16; }();
  // dQ reduce-add boxes: 16 warps -> one [32 rows x 16 cols] fp32 box per warp (SWIZZLE_64B); 8 warps -> [128 x 32] (128B)
  if (ncw == 16) s = make_tmap_2d_ex(&tmDQ, a->dq_acc, 4, (uint64_t)f->H * 64, (uint64_t)f->B * f->Lq, a->ld_dq, 16, 32, 64);
  else s = make_tmap_2d_ex(&tmDQ, a->dq_acc, 4, (uint64_t)f->H * 64, (uint64_t)f->B * f->Lq, a->ld_dq, 32, 128, 128);
  if (s != VC_OK) return s;
  if ((s = make_tmap_3d(&tmQ, f->q, f->ldq, f->Lq, f->B, f->ldq, (uint64_t)f->Lq * f->ldq, 64, kBT)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmK, f->k, f->ldk, f->Lk, f->B, f->ldk, (uint64_t)f->Lk * f->ldk, 64, kBT)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmV, f->v, f->ldv, f->Lk, f->B, f->ldv, (uint64_t)f->Lk * f->ldv, 64, kBT)) != VC_OK) return s;
  if ((s = make_tmap_3d(&tmDO, a->dout, a->ld_do, f->Lq, f->B, a->ld_do, (uint64_t)f->Lq * a->ld_do, 64, kBT)) != VC_OK) return s;
  AttnBwdParams p;
  p.B = f->B; p.H = f->H; p.Lq = f->Lq; p.Lk = f->Lk;
  p.q_col = f->q_col; p.k_col = f->k_col; p.v_col = f->v_col; p.do_col = a->do_col;
  p.lse2 = f->lse2; p.delta = a->delta; p.bias_rel = f->bias_rel; p.bucket_lut = a->bucket_lut; p.kmask = f->kmask;
  p.causal = f->causal; p.scale = f->scale; p.scale_log2e = f->scale * kBLog2e; p.inv_scale = 1.0f / f->scale;
  p.dq_acc = a->dq_acc; p.ld_dq = a->ld_dq;
  p.dk = (__nv_bfloat16*)a->dk; p.ld_dk = a->ld_dk; p.dk_col = a->dk_col;
  p.dv = (__nv_bfloat16*)a->dv; p.ld_dv = a->ld_dv; p.dv_col = a->dv_col;
  p.dbias_rel = a->dbias_rel;
  p.drop_seed = f->drop_seed; p.drop_p16 = f->drop_p16; p.drop_salt = drop_salt_ptr();
  p.q_like_k = f->q_like_k;
  p.trace = debug_trace_ptr();
  VC_CHECK(!f->q_like_k || (f->Lq == f->Lk && f->kmask && !f->causal), "vc_attn_bwd: q_like_k needs non-causal self-attention with a key mask");
  static PerDeviceOnce attr;
  if (attr.need()) {
    VC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnBwdSmem));
    VC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnBwdSmem));
  }
  dim3 grid((f->Lk + kBT - 1) / kBT, f->H, f->B);
  if (ncw == 8) VC_CUDA(launch_kernel(attn_bwd_kernel<8>, grid, dim3(320), (size_t)kAttnBwdSmem, st, tmQ, tmK, tmV, tmDO, tmDQ, p));
  else VC_CUDA(launch_kernel(attn_bwd_kernel<16>, grid, dim3(576), (size_t)kAttnBwdSmem, st, tmQ, tmK, tmV, tmDO, tmDQ, p));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
