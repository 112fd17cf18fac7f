// Shared pieces of the tcgen05 GEMM kernels (1-CTA gemm.cu, 2-CTA gemm2.cu): parameters and the fused epilogue.
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

constexpr int BM = 128;
constexpr int BK = 64;

struct GemmParams {
  int M, N, K;
  int m_blocks, n_blocks, splits, kb_per_split, num_kb;
  // epilogue
  void* out;             // bf16 or fp32 [M][ldo]
  long long ldo;
  int out_fp32;
  int atomic;            // fp32 out: atomicAdd instead of store
  const float* bias;     // [N] or null
  const float* residual; // fp32 [M][ldr] or null  (may alias out)
  long long ldr;
  int act;               // 0 none, 1 relu, 2 gelu(erf), 3 mul relu'(aux), 4 mul gelu'(aux)
  __nv_bfloat16* pre_out;  // bf16 [M][ldo] pre-activation copy (act=2) or null
  const __nv_bfloat16* aux;  // bf16 [M][ld_aux] for act 3/4
  long long ld_aux;
  float alpha;
  const float* alpha_dev;  // optional device scalar multiplied into alpha
  uint32_t drop_seed, drop_p16;  // dropout applied after the activation, before the residual add (p16 = 0: off)
  const uint32_t* drop_salt;     // optional device salt XORed into drop_seed
  int tma_epi;                   // 1: epilogue I/O staged through shared memory and moved by TMA (see below)
  int aux_tma;                   // 1: the activation-backward operand `aux` arrives by TMA too (EpiMaps::resid maps it)
  // fused LM head + label-smoothed cross entropy (act 5: per-row partial statistics, act 6: d loss / d logits)
  const long long* ce_labels;
  float* ce_stats;               // [M][ce_slots][3]
  float* ce_zy;                  // [M]
  const float* ce_lse;           // [M]
  const float* ce_nvalid;        // device scalar
  float ce_smoothing;
  int ce_slots;                  // 2 * n_blocks
  int dbg;                       // measurement only (VIDCHAP_GEMM_DBG=1): the epilogue skips its chunks (tools/time_gemm_epi.py)
  int epi_preset;                // index into epi_preset_mask(): which compiled epilogue the pair kernel runs
};

constexpr int ACT_CE_STATS = 5, ACT_CE_GRAD = 6;

// act 5: one 32-column chunk of one row folded into the running (max, sum exp, sum z) of the thread's row.
__device__ __forceinline__ void ce_stats_chunk(const GemmParams& p, uint32_t taddr, int row, bool row_ok, int col0, float alpha,
                                               float& m, float& s, float& t) {
  float v[32];
  tmem_ld32(taddr, v);
  const long long y = row_ok ? p.ce_labels[row] : -1;
  tmem_ld_wait();
  const int ncols = min(32, p.N - col0);
  if (!row_ok || ncols <= 0) return;
  constexpr float kL2e = 1.4426950408889634f;
  float cm = -INFINITY, cs = 0.f, ct = 0.f;
  if (ncols == 32) {          // whole chunk in range (all but the last column block of a ragged vocabulary)
#pragma unroll
    for (int j = 0; j < 32; ++j) { v[j] *= alpha; cm = fmaxf(cm, v[j]); }
    const float m_new = fmaxf(m, cm);
    const float neg = -m_new * kL2e;
#pragma unroll
    for (int j = 0; j < 32; ++j) { cs += fast_exp2(fmaf(v[j], kL2e, neg)); ct += v[j]; }
    s = s * fast_exp2((m - m_new) * kL2e) + cs;     // first chunk: m = -inf -> factor 0
    m = m_new;
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] *= alpha;
      if (j < ncols) cm = fmaxf(cm, v[j]);
    }
    const float m_new = fmaxf(m, cm);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols) { cs += __expf(v[j] - m_new); ct += v[j]; }
    }
    s = s * __expf(m - m_new) + cs;
    m = m_new;
  }
  t += ct;
  if (y >= col0 && y < col0 + ncols) {
    float zy = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) zy = (col0 + j == y) ? v[j] : zy;
    p.ce_zy[row] = zy;
  }
}

// act 6: logits chunk -> d loss / d logits (in place)
__device__ __forceinline__ void ce_grad_chunk(const GemmParams& p, float* v, int row, bool row_ok, int col0) {
  const long long y = row_ok ? p.ce_labels[row] : -100;
  const float lse = row_ok ? p.ce_lse[row] : 0.f;
  const float inv_nv = 1.0f / __ldg(p.ce_nvalid);
  const float sm = p.ce_smoothing / (float)p.N;
  if (y < 0) {   // ignore_index
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    return;
  }
  constexpr float kL2e = 1.4426950408889634f;
  const float neg = -lse * kL2e, off = -sm * inv_nv;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = fmaf(fast_exp2(fmaf(v[j], kL2e, neg)), inv_nv, off);
  if (y >= col0 && y < col0 + 32) {      // the one label column of this row, if it falls into this chunk
    const float hot = (1.0f - p.ce_smoothing) * inv_nv;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (col0 + j == y) ? v[j] - hot : v[j];
  }
}

// Tensor maps of the epilogue's global operands (32 x 32 boxes; fp32: SWIZZLE_128B, bf16: SWIZZLE_64B).
struct EpiMaps {
  CUtensorMap out, pre, resid;
};

constexpr int kEpiStageBytes = 4096;   // per epilogue warp: 32 rows x 128 B
constexpr int kEpiWarps = 8;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// One 32-column chunk of one accumulator row: TMEM -> registers -> fused epilogue -> global.
// Epilogue order: *alpha, +bias, (pre_out copy), activation / activation-backward, dropout, +residual, store.
// Side inputs (residual, aux) are requested BEFORE waiting for the TMEM load so their latency overlaps it.
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, uint32_t taddr, int row, bool row_ok, int col0,
                                                    float alpha) {
          float v[32];
          tmem_ld32(taddr, v);
          const int ncols = min(32, p.N - col0);
          const bool full = row_ok && ncols == 32;  // fast path: whole 32-column chunk in range
          float4 rs[8];
          uint4 ax[4];
          if (full) {
            if (p.residual) {
              const float4* rp = reinterpret_cast<const float4*>(p.residual + (long long)row * p.ldr + col0);
  #pragma unroll
              for (int j = 0; j < 8; ++j) rs[j] = rp[j];
            }
            if (p.act == 3 || p.act == 4) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.aux + (long long)row * p.ld_aux + col0);
  #pragma unroll
              for (int j = 0; j < 4; ++j) ax[j] = __ldg(ap + j);
            }
          }
          tmem_ld_wait();
          if (!row_ok || ncols <= 0) return;
  #pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= alpha;
          if (p.bias) {
            if (full) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
  #pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(bp + j);
                v[j * 4] += b4.x; v[j * 4 + 1] += b4.y; v[j * 4 + 2] += b4.z; v[j * 4 + 3] += b4.w;
              }
            } else {
  #pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
            }
          }
          if (p.act == 2 && p.pre_out) {
            __nv_bfloat16* dstp = p.pre_out + (long long)row * p.ldo + col0;
            if (full) {
              uint4* dst = reinterpret_cast<uint4*>(dstp);
  #pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                    pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
            } else {
  #pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncols) dstp[j] = __float2bfloat16(v[j]);
            }
          }
          if (p.act == 1) {
  #pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          } else if (p.act == 2) {
  #pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          } else if (p.act == ACT_CE_GRAD) {
            ce_grad_chunk(p, v, row, row_ok, col0);
          } else if (p.act >= 3) {
            if (!full) {  // ragged tail: gather the aux values element-wise
              const __nv_bfloat16* axp = p.aux + (long long)row * p.ld_aux + col0;
              uint32_t w[16];
  #pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float lo = e * 2 < ncols ? __bfloat162float(axp[e * 2]) : 0.f;
                const float hi = e * 2 + 1 < ncols ? __bfloat162float(axp[e * 2 + 1]) : 0.f;
                w[e] = pack_bf16x2(lo, hi);
              }
  #pragma unroll
              for (int j = 0; j < 4; ++j) ax[j] = make_uint4(w[j * 4], w[j * 4 + 1], w[j * 4 + 2], w[j * 4 + 3]);
            }
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {ax[j].x, ax[j].y, ax[j].z, ax[j].w};
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
                if (p.act == 3) {
                  v[j * 8 + e * 2] = lo > 0.0f ? v[j * 8 + e * 2] : 0.0f;
                  v[j * 8 + e * 2 + 1] = hi > 0.0f ? v[j * 8 + e * 2 + 1] : 0.0f;
                } else {
                  v[j * 8 + e * 2] *= gelu_erf_grad(lo);
                  v[j * 8 + e * 2 + 1] *= gelu_erf_grad(hi);
                }
              }
            }
          }
          if (p.drop_p16) {
            drop_apply<32>(v, drop_row_key(drop_salted(p.drop_seed, p.drop_salt), (unsigned long long)row), p.drop_p16, (uint32_t)col0,
                           drop_scale(p.drop_p16));
          }
          if (p.residual) {
            if (full) {
  #pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[j * 4 + 0] += rs[j].x; v[j * 4 + 1] += rs[j].y; v[j * 4 + 2] += rs[j].z; v[j * 4 + 3] += rs[j].w;
              }
            } else {
              const float* rsp = p.residual + (long long)row * p.ldr + col0;
  #pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += rsp[j];
            }
          }
          if (p.out_fp32) {
            float* dstf = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
            if (full) {
              float4* dst = reinterpret_cast<float4*>(dstf);
              if (p.atomic) {
  #pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(dst + j, make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]));
              } else {
  #pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
              }
            } else {
  #pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < ncols) {
                  if (p.atomic) atomicAdd(dstf + j, v[j]); else dstf[j] = v[j];
                }
              }
            }
          } else {
            __nv_bfloat16* dstb = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col0;
            if (full) {
              uint4* dst = reinterpret_cast<uint4*>(dstb);
  #pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                    pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
            } else {
  #pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncols) dstb[j] = __float2bfloat16(v[j]);
            }
          }
}

// ---- TMA epilogue.
// Why: with lane <-> accumulator row (the TMEM layout), direct global stores/loads touch 32 different rows per warp
// instruction in 16-byte pieces; ncu showed the L1TEX pipe ~70 % busy with that traffic while the tensor pipe sat at
// 46 % (profiles/r01_ncu_prof_gemm1_fwd_r01.txt) — the same L1TEX pipe also carries the TMA fills of the mainloop.
// Here every warp stages its 32 x 32 block in a swizzled shared-memory tile: the residual block arrives by TMA load,
// results leave by TMA store (or TMA reduce-add for split-K / accumulate), i.e. whole 128-byte lines on both sides.
__device__ __forceinline__ uint32_t epi_off_f32(int t, int g) { return t * 128 + ((g ^ (t & 7)) << 4); }        // SWIZZLE_128B
__device__ __forceinline__ uint32_t epi_off_bf16(int t, int g) { return t * 64 + ((g ^ ((t >> 1) & 3)) << 4); }  // SWIZZLE_64B

// Compile-time feature set of an epilogue instantiation: a bit that is clear removes the code of that feature (and the
// constant-bank loads and uniform branches that select it at run time); a bit that is set leaves the run-time test in.
// The pair kernel carries a handful of presets for the train step's hot GEMMs next to the all-features instantiation
// (epi_preset_for() below): ncu had the epilogue warps of the generic code issue-bound at ~340 instructions per 32 x 32
// chunk, longer than the K = 768 mainloop of the same tile (profiles/r02_gemm_epilogue.md).
enum : uint32_t {
  EF_BIAS = 1u, EF_RELU = 2u, EF_GELU = 4u, EF_ACT_BWD = 8u, EF_CE_GRAD = 16u, EF_DROP = 32u, EF_RESID = 64u,
  EF_F32 = 128u, EF_BF16 = 256u, EF_ATOMIC = 512u, EF_ALL = 1023u
};
constexpr int kEpiPresets = 6;
__host__ __device__ constexpr uint32_t epi_preset_mask(int i) {
  return i == 0 ? (EF_BF16 | EF_DROP)                            // Q/K/V projections, plain dgrads
       : i == 1 ? (EF_BF16 | EF_RELU | EF_DROP)                  // wi + ReLU (+ dropout)
       : i == 2 ? (EF_F32 | EF_RESID | EF_DROP)                  // o / wo projections into the fp32 residual stream
       : i == 3 ? (EF_BF16 | EF_ACT_BWD | EF_DROP)               // dgrad through ReLU / GELU (saved pre-activation)
       : i == 4 ? (EF_F32 | EF_ATOMIC)                           // wgrad (split-K reduce-add) and plain fp32 outputs
       : EF_ALL;
}

template <uint32_t F>
__device__ __forceinline__ void gemm_epilogue_chunk_tma(const GemmParams& p, const EpiMaps& em, uint32_t taddr, int row0,
                                                        int lane, int col0, float alpha, uint8_t* stage, uint64_t* wbar,
                                                        uint32_t& wphase) {
  float v[32];
  tmem_ld32(taddr, v);
  const int row = row0 + lane;
  const bool in_rows = row < p.M;
  const bool out_f32 = !(F & EF_BF16) ? true : !(F & EF_F32) ? false : p.out_fp32 != 0;
  const bool has_resid = (F & EF_RESID) && p.residual;
  const bool act_bwd = (F & EF_ACT_BWD) && (p.act == 3 || p.act == 4);
  const bool aux_tma = act_bwd && p.aux_tma;
  const bool two_stores = (F & EF_GELU) && p.act == 2 && p.pre_out;  // the staging tile carries the pre-activation copy first
  // dropout scale folded into alpha where everything between the two is linear or ReLU
  constexpr bool kFoldDrop = (F & EF_DROP) && !(F & (EF_BIAS | EF_GELU | EF_CE_GRAD));
  const uint32_t drop_p16 = (F & EF_DROP) ? p.drop_p16 : 0u;
  if (kFoldDrop && drop_p16) alpha *= drop_scale(drop_p16);
  if (lane == 0) {
    bulk_wait_read0();  // the previous chunk's store has finished reading this staging tile
    if (has_resid && !two_stores) {
      mbar_arrive_expect_tx(wbar, 4096);
      tma_load_2d(stage, &em.resid, wbar, col0, row0);
    } else if (aux_tma) {   // saved pre-activation block (bf16 32 x 32): whole lines instead of 32 scattered rows
      mbar_arrive_expect_tx(wbar, 2048);
      tma_load_2d(stage, &em.resid, wbar, col0, row0);
    }
  }
  __syncwarp();
  uint4 ax[4];
  if (act_bwd && !aux_tma) {
    if (in_rows && col0 + 32 <= p.N) {
      const uint4* ap = reinterpret_cast<const uint4*>(p.aux + (long long)row * p.ld_aux + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) ax[j] = __ldg(ap + j);
    } else {
      const __nv_bfloat16* axp = p.aux + (long long)row * p.ld_aux + col0;
      uint32_t w[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float lo = (in_rows && col0 + e * 2 < p.N) ? __bfloat162float(axp[e * 2]) : 0.f;
        const float hi = (in_rows && col0 + e * 2 + 1 < p.N) ? __bfloat162float(axp[e * 2 + 1]) : 0.f;
        w[e] = pack_bf16x2(lo, hi);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) ax[j] = make_uint4(w[j * 4], w[j * 4 + 1], w[j * 4 + 2], w[j * 4 + 3]);
    }
  }
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= alpha;
  if ((F & EF_BIAS) && p.bias) {
    if (col0 + 32 <= p.N) {
      const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = __ldg(bp + j);
        v[j * 4] += b4.x; v[j * 4 + 1] += b4.y; v[j * 4 + 2] += b4.z; v[j * 4 + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (two_stores) {  // pre-activation copy (bf16) leaves first
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(stage + epi_off_bf16(lane, g)) =
          make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                     pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&em.pre, stage, col0, row0);
      bulk_commit();
      bulk_wait_read0();
      if (has_resid) {  // only now may the residual block land in the staging tile
        mbar_arrive_expect_tx(wbar, 4096);
        tma_load_2d(stage, &em.resid, wbar, col0, row0);
      }
    }
    __syncwarp();
  }
  if ((F & EF_RELU) && p.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if ((F & EF_GELU) && p.act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if ((F & EF_CE_GRAD) && p.act == ACT_CE_GRAD) {
    ce_grad_chunk(p, v, row, in_rows, col0);
  } else if (act_bwd) {
    if (aux_tma) {
      mbar_wait(wbar, wphase);
      wphase ^= 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) ax[j] = *reinterpret_cast<const uint4*>(stage + epi_off_bf16(lane, j));
      __syncwarp();   // every lane has read its row before the tile is overwritten with the result
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w[4] = {ax[j].x, ax[j].y, ax[j].z, ax[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
        if (p.act == 3) {
          v[j * 8 + e * 2] = lo > 0.0f ? v[j * 8 + e * 2] : 0.0f;
          v[j * 8 + e * 2 + 1] = hi > 0.0f ? v[j * 8 + e * 2 + 1] : 0.0f;
        } else {
          v[j * 8 + e * 2] *= gelu_erf_grad(lo);
          v[j * 8 + e * 2 + 1] *= gelu_erf_grad(hi);
        }
      }
    }
  }
  if (drop_p16) {
    const uint32_t key = drop_row_key(drop_salted(p.drop_seed, p.drop_salt), (unsigned long long)row);
    if (kFoldDrop) drop_select<32>(v, key, drop_p16, (uint32_t)col0);
    else drop_apply<32>(v, key, drop_p16, (uint32_t)col0, drop_scale(drop_p16));
  }
  if (has_resid) {
    mbar_wait(wbar, wphase);
    wphase ^= 1;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 r = *reinterpret_cast<const float4*>(stage + epi_off_f32(lane, g));
      v[g * 4 + 0] += r.x; v[g * 4 + 1] += r.y; v[g * 4 + 2] += r.z; v[g * 4 + 3] += r.w;
    }
  }
  if (out_f32) {
#pragma unroll
    for (int g = 0; g < 8; ++g)
      *reinterpret_cast<float4*>(stage + epi_off_f32(lane, g)) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(stage + epi_off_bf16(lane, g)) =
          make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                     pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if ((F & EF_ATOMIC) && p.atomic) tma_reduce_add_2d(&em.out, stage, col0, row0);
    else tma_store_2d(&em.out, stage, col0, row0);
    bulk_commit();
  }
}

// Feature bits one call needs, and the first (narrowest) preset that covers them.
inline uint32_t epi_features(const vc_gemm_args* a) {
  uint32_t f = a->out_fp32 ? EF_F32 : EF_BF16;
  if (a->bias) f |= EF_BIAS;
  if (a->act == 1) f |= EF_RELU;
  if (a->act == 2) f |= EF_GELU;
  if (a->act == 3 || a->act == 4) f |= EF_ACT_BWD;
  if (a->act == ACT_CE_GRAD) f |= EF_CE_GRAD;
  if (a->drop_p16) f |= EF_DROP;
  if (a->residual) f |= EF_RESID;
  if (a->atomic) f |= EF_ATOMIC;
  return f;
}
inline int epi_preset_for(const vc_gemm_args* a) {
  const uint32_t need = epi_features(a);
  for (int i = 0; i < kEpiPresets; ++i)
    if ((need & ~epi_preset_mask(i)) == 0) return i;
  return kEpiPresets - 1;
}

// act-backward GEMMs (no residual): the `resid` map slot carries the bf16 `aux` matrix instead.
inline bool epi_aux_by_tma(const vc_gemm_args* a) {
  return (a->act == 3 || a->act == 4) && a->aux && !a->residual && !(a->act == 2 && a->pre_out) && a->ld_aux % 8 == 0 &&
         ((uintptr_t)a->aux & 15) == 0;
}

// Host side: the three epilogue maps (unused ones alias `out`).
inline int make_epi_maps(EpiMaps* em, const vc_gemm_args* a) {
  const int oe = a->out_fp32 ? 4 : 2;
  int s = make_tmap_2d_ex(&em->out, a->out, oe, a->N, a->M, a->ldo, 32, 32, a->out_fp32 ? 128 : 64);
  if (s != VC_OK) return s;
  em->pre = em->out;
  em->resid = em->out;
  if (a->pre_out && (s = make_tmap_2d_ex(&em->pre, a->pre_out, 2, a->N, a->M, a->ldo, 32, 32, 64)) != VC_OK) return s;
  if (a->residual && (s = make_tmap_2d_ex(&em->resid, a->residual, 4, a->N, a->M, a->ldr, 32, 32, 128)) != VC_OK) return s;
  if (epi_aux_by_tma(a) && (s = make_tmap_2d_ex(&em->resid, a->aux, 2, a->N, a->M, a->ld_aux, 32, 32, 64)) != VC_OK) return s;
  return VC_OK;
}

}  // namespace vc
