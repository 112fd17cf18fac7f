// bf16 GEMM on tcgen05 tensor cores (sm_100a): C[M,N] = op(A)[M,K] * op(B)[N,K]^T, fp32 accumulate in TMEM.
//
// Replaces every nn.Linear / F.linear on the Vid2Seq path (reference: model/modeling_t5.py:305,310,528-536,581,1714;
// model/vit.py:17-20,41,53) and their autograd backward (dgrad / wgrad), which the reference runs as cuBLAS SGEMM.
//
// Layout in HBM: all operands bf16 row-major.
//   A "K-major"  : stored [M][K]  (activations in forward, dY in dgrad)
//   A "MN-major" : stored [K][M]  (dY^T in wgrad: reduction over tokens, tokens are the rows)
//   B "K-major"  : stored [N][K]  (nn.Linear weight in forward)
//   B "MN-major" : stored [K][N]  (weight in dgrad, activations in wgrad)
// Kernel: persistent, one CTA per SM, 384 threads:
//   warp 0 lane 0 : TMA producer      (cp.async.bulk.tensor, 128B-swizzled 64-wide boxes -> 4..8 stage smem ring)
//   warp 1 lane 0 : tcgen05.mma issuer (UMMA 128 x BN x 16, cta_group::1), commits to mbarriers
//   warp 2        : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4..11   : epilogue (2 warps per TMEM lane quarter, one per column half): tcgen05.ld TMEM->registers, fused bias / activation / activation-backward / residual /
//                   alpha, bf16 or fp32 store, or fp32 atomic accumulate (split-K for wgrad).
// The accumulator is double-buffered in TMEM so tile i's epilogue overlaps tile i+1's MMAs.
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
};

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytes = (BM + BN) * BK * 2;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // 512 / 256 / 128
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kABytes = BM * BK * 2;
  constexpr int kBBytes = BN * BK * 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.m_blocks * p.n_blocks;
  const int num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mn = t % tiles_mn, split = t / tiles_mn;
      const int m0 = (mn % p.m_blocks) * BM, n0 = (mn / p.m_blocks) * BN;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], kABytes + kBBytes);
        uint8_t* sa = smem_a + stage * kABytes;
        uint8_t* sb = smem_b + stage * kBBytes;
        const int k0 = kb * BK;
        if (!A_MN) {
          tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);  // box 64(k) x 128(m)
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, &full_bar[stage], m0 + j * 64, k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);  // box 64(k) x BN(n)
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[stage], n0 + j * 64, k0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int split = t / tiles_mn;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
        const uint32_t b_addr = smem_u32(smem_b + stage * kBBytes);
        // K-major: SBO = 1024 (8 rows x 128 B), LBO unused.  MN-major: SBO = 1024 (8 k-rows), LBO = 64-wide atom stride.
        const uint64_t adesc = make_smem_desc_sw128(a_addr, A_MN ? BK * 128 : 0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(b_addr, B_MN ? BK * 128 : 0, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = adesc + (uint64_t)((A_MN ? k * 2048 : k * 32) >> 4);
          const uint64_t bd = bdesc + (uint64_t)((B_MN ? k * 2048 : k * 32) >> 4);
          tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(&tmem_full[acc]);  // accumulator complete
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps =====================
    // Warp w reads TMEM lane quarter (w & 3) and the column half (w - 4) / 4 of the tile.  Side inputs (residual, aux)
    // are requested BEFORE waiting for the TMEM load so their L2/HBM latency overlaps it; with the accumulator double
    // buffered the whole epilogue of tile i overlaps the MMAs of tile i+1.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kChunks = BN / 64;  // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mn = t % tiles_mn;
      const int m0 = (mn % p.m_blocks) * BM, n0 = (mn / p.m_blocks) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int cc = half * kChunks + c;  // chunk index inside the tile
        const int col0 = n0 + cc * 32;
        float v[32];
        tmem_ld32(tmem_base + acc * BN + cc * 32 + ((uint32_t)(q * 32) << 16), v);
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
          if (p.act >= 3) {
            const uint4* ap = reinterpret_cast<const uint4*>(p.aux + (long long)row * p.ld_aux + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) ax[j] = __ldg(ap + j);
          }
        }
        tmem_ld_wait();
        if (!row_ok || ncols <= 0) continue;
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
          const float sc = drop_scale(p.drop_p16);
          const unsigned long long base = (unsigned long long)row * (unsigned long long)p.N + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = drop_keep(p.drop_seed, p.drop_p16, base + j) ? v[j] * sc : 0.0f;
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
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  kern<<<grid, 384, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc

using namespace vc;

extern "C" int vc_gemm_bf16(const vc_gemm_args* a, void* stream) {
  VC_CHECK(a != nullptr, "vc_gemm_bf16: null args");
  VC_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "vc_gemm_bf16: bad dims M=%d N=%d K=%d", a->M, a->N, a->K);
  VC_CHECK(a->lda % 8 == 0 && a->ldb % 8 == 0, "vc_gemm_bf16: lda/ldb must be multiples of 8 elements (16 B)");
  VC_CHECK(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->B & 15) == 0 && ((uintptr_t)a->out & 15) == 0,
           "vc_gemm_bf16: A/B/out must be 16-byte aligned");
  VC_CHECK(a->ldo % (a->out_fp32 ? 4 : 8) == 0, "vc_gemm_bf16: ldo alignment");
  VC_CHECK(!a->atomic || a->out_fp32, "vc_gemm_bf16: atomic accumulate needs fp32 out");
  VC_CHECK(a->act >= 0 && a->act <= 4, "vc_gemm_bf16: bad act %d", a->act);
  VC_CHECK((a->act != 3 && a->act != 4) || a->aux, "vc_gemm_bf16: act %d needs aux", a->act);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  int BN = a->tile_n;
  if (BN == 0) {
    // Pick the widest N tile that still yields >= ~1 wave of CTAs.
    const int mb = (a->M + BM - 1) / BM;
    const int sms = num_sms();
    BN = 256;
    while (BN > 64 && (long long)mb * ((a->N + BN - 1) / BN) * (a->splits > 0 ? a->splits : 1) < sms) BN >>= 1;
  }
  VC_CHECK(BN == 64 || BN == 128 || BN == 256, "vc_gemm_bf16: tile_n must be 0/64/128/256");

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_blocks = (a->M + BM - 1) / BM;
  p.n_blocks = (a->N + BN - 1) / BN;
  p.num_kb = (a->K + BK - 1) / BK;
  int splits = a->splits > 0 ? a->splits : 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  p.splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  VC_CHECK(p.splits == 1 || a->atomic, "vc_gemm_bf16: split-K needs atomic fp32 accumulate");
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = a->out_fp32; p.atomic = a->atomic;
  p.bias = a->bias; p.residual = a->residual; p.ldr = a->ldr; p.act = a->act;
  p.pre_out = reinterpret_cast<__nv_bfloat16*>(a->pre_out);
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); p.ld_aux = a->ld_aux;
  p.alpha = a->alpha;
  p.alpha_dev = a->alpha_dev;
  p.drop_seed = a->drop_seed; p.drop_p16 = a->drop_p16;

  CUtensorMap tmA, tmB;
  int s;
  if (!a->a_mn_major) s = make_tmap_2d(&tmA, a->A, a->K, a->M, a->lda, 64, BM);
  else                s = make_tmap_2d(&tmA, a->A, a->M, a->K, a->lda, 64, BK);
  if (s != VC_OK) return s;
  if (!a->b_mn_major) s = make_tmap_2d(&tmB, a->B, a->K, a->N, a->ldb, 64, BN);
  else                s = make_tmap_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, BK);
  if (s != VC_OK) return s;

  const int num_tiles = p.m_blocks * p.n_blocks * p.splits;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();

#define VC_DISPATCH(BN_)                                                                              \
  if (BN == BN_) {                                                                                    \
    if (!a->a_mn_major && !a->b_mn_major) return launch_gemm<BN_, false, false>(tmA, tmB, p, grid, st); \
    if (!a->a_mn_major && a->b_mn_major) return launch_gemm<BN_, false, true>(tmA, tmB, p, grid, st);   \
    if (a->a_mn_major && !a->b_mn_major) return launch_gemm<BN_, true, false>(tmA, tmB, p, grid, st);   \
    return launch_gemm<BN_, true, true>(tmA, tmB, p, grid, st);                                        \
  }
  VC_DISPATCH(256)
  VC_DISPATCH(128)
  VC_DISPATCH(64)
#undef VC_DISPATCH
  return VC_ERR_INVALID;
}
