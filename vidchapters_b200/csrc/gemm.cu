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
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"

namespace vc {

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytes = (BM + BN) * BK * 2;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // 512 / 256 / 128
};

template <int BN, bool A_MN, bool B_MN, bool TMA_EPI>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ EpiMaps em, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kABytes = BM * BK * 2;
  constexpr int kBBytes = BN * BK * 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* epi_stage = smem + kStages * Cfg::kStageBytes;   // 8 x 4 KB staging tiles of the TMA epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + kEpiWarps * kEpiStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* epi_bar = tmem_empty + 2;   // [8] one per epilogue warp (residual tile landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + kEpiWarps);

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
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // prologue above overlapped the previous kernel's tail; from here on global memory is read
  pdl_trigger();
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
    uint32_t acc_phase = 0, epi_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mn = t % tiles_mn;
      const int m0 = (mn % p.m_blocks) * BM, n0 = (mn / p.m_blocks) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
      if (p.act == ACT_CE_STATS) {
        // fused cross entropy, statistics pass: nothing is stored; this warp's half of the tile's columns is folded into
        // (max, sum exp, sum z) of its rows and written to the row's slot of this (column block, half)
        float cm = -INFINITY, cs = 0.f, ct = 0.f;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          const int cc = half * kChunks + c;
          ce_stats_chunk(p, tmem_base + acc * BN + cc * 32 + ((uint32_t)(q * 32) << 16), row, row_ok, n0 + cc * 32, alpha, cm, cs, ct);
        }
        if (row_ok) {
          float* dst = p.ce_stats + ((long long)row * p.ce_slots + (n0 / BN) * 2 + half) * 3;
          dst[0] = cm; dst[1] = cs; dst[2] = ct;
        }
      } else
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int cc = half * kChunks + c;  // chunk index inside the tile
        const uint32_t taddr = tmem_base + acc * BN + cc * 32 + ((uint32_t)(q * 32) << 16);
        if constexpr (TMA_EPI)
          gemm_epilogue_chunk_tma<EF_ALL>(p, em, taddr, m0 + q * 32, lane, n0 + cc * 32, alpha, epi_stage + (warp - 4) * kEpiStageBytes,
                                  &epi_bar[warp - 4], epi_phase);
        else
          gemm_epilogue_chunk(p, taddr, row, row_ok, n0 + cc * 32, alpha);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMA_EPI && lane == 0) bulk_wait_all();  // staging tiles must outlive the TMA stores that read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, bool TMA_EPI>
static int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, const GemmParams& p, int grid,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, TMA_EPI>;
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  VC_CUDA(launch_kernel(kern, dim3(grid), dim3(384), Cfg::kSmemBytes, st, tmA, tmB, em, p));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, const GemmParams& p, int grid,
                       cudaStream_t st) {
  return p.tma_epi ? launch_gemm_t<BN, A_MN, B_MN, true>(tmA, tmB, em, p, grid, st)
                   : launch_gemm_t<BN, A_MN, B_MN, false>(tmA, tmB, em, p, grid, st);
}

}  // namespace vc

namespace vc {
int launch_gemm_pair(const vc_gemm_args* a, int BN, const EpiMaps& em, const GemmParams& p, cudaStream_t st);  // gemm2.cu
static int tma_epi_mode() {  // VIDCHAP_GEMM_TMA_EPI=0 selects the direct-store epilogue (A/B testing); default on
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VIDCHAP_GEMM_TMA_EPI");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode;
}
static int pair_mode() {  // VIDCHAP_GEMM_PAIR=0 disables the 2-CTA kernel (A/B testing); default on
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VIDCHAP_GEMM_PAIR");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode;
}
}  // namespace vc

using namespace vc;

extern "C" int vc_gemm_bf16(const vc_gemm_args* a, void* stream) {
  VC_CHECK(a != nullptr, "vc_gemm_bf16: null args");
  VC_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "vc_gemm_bf16: bad dims M=%d N=%d K=%d", a->M, a->N, a->K);
  VC_CHECK(a->lda % 8 == 0 && a->ldb % 8 == 0, "vc_gemm_bf16: lda/ldb must be multiples of 8 elements (16 B)");
  VC_CHECK(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->B & 15) == 0 && ((uintptr_t)a->out & 15) == 0,
           "vc_gemm_bf16: A/B/out must be 16-byte aligned");
  VC_CHECK(a->out || a->act == 5, "vc_gemm_bf16: null out");
  VC_CHECK(a->ldo % (a->out_fp32 ? 4 : 8) == 0, "vc_gemm_bf16: ldo alignment");
  VC_CHECK(!a->atomic || a->out_fp32, "vc_gemm_bf16: atomic accumulate needs fp32 out");
  VC_CHECK(a->act >= 0 && a->act <= 6, "vc_gemm_bf16: bad act %d", a->act);
  if (a->act == ACT_CE_STATS || a->act == ACT_CE_GRAD) {
    VC_CHECK(a->ce_labels && a->ce_nvalid, "vc_gemm_bf16: fused cross entropy needs ce_labels / ce_nvalid");
    VC_CHECK(!a->bias && !a->residual && a->drop_p16 == 0 && (a->splits <= 1) && !a->atomic && !a->a_mn_major,
             "vc_gemm_bf16: fused cross entropy is a plain alpha*A.B^T epilogue");
    if (a->act == ACT_CE_STATS) VC_CHECK(a->ce_stats && a->ce_zy && (a->tile_n == 128 || a->tile_n == 256),
                                         "vc_gemm_bf16: act 5 needs ce_stats, ce_zy and an explicit tile_n (128/256)");
    else VC_CHECK(a->ce_lse && !a->out_fp32 && a->out, "vc_gemm_bf16: act 6 needs ce_lse and a bf16 out");
  }
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
  VC_CHECK(p.splits == 1 || (!a->bias && !a->residual && a->act == 0 && a->drop_p16 == 0),
           "vc_gemm_bf16: split-K supports the plain alpha*A.B accumulate epilogue only");
  p.out = a->out; p.ldo = a->ldo; p.out_fp32 = a->out_fp32; p.atomic = a->atomic;
  p.bias = a->bias; p.residual = a->residual; p.ldr = a->ldr; p.act = a->act;
  p.pre_out = reinterpret_cast<__nv_bfloat16*>(a->pre_out);
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); p.ld_aux = a->ld_aux;
  p.alpha = a->alpha;
  p.alpha_dev = a->alpha_dev;
  p.drop_seed = a->drop_seed; p.drop_p16 = a->drop_p16; p.drop_salt = drop_salt_ptr();
  p.ce_labels = reinterpret_cast<const long long*>(a->ce_labels); p.ce_stats = a->ce_stats; p.ce_zy = a->ce_zy;
  p.ce_lse = a->ce_lse; p.ce_nvalid = a->ce_nvalid; p.ce_smoothing = a->ce_smoothing;
  p.ce_slots = 2 * p.n_blocks;
  static const int dbg = [] { const char* e = getenv("VIDCHAP_GEMM_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  p.epi_preset = epi_preset_for(a);
  p.tma_epi = tma_epi_mode() && (a->ldr % 4 == 0) && a->act != ACT_CE_STATS;
  p.aux_tma = p.tma_epi && epi_aux_by_tma(a);
  EpiMaps em;
  if (p.tma_epi) {
    int se = make_epi_maps(&em, a);
    if (se != VC_OK) return se;
  } else {
    memset(&em, 0, sizeof(em));
  }

  // CTA-pair kernel (256 x BN tiles) whenever the problem still fills the machine with pairs
  // VIDCHAP_GEMM_PAIR_MIN_M: smallest M that takes the CTA-pair kernel (A/B switch; smaller problems are launch /
  // prologue bound and the pair's cluster launch and two cluster barriers are pure overhead there)
  static const int pair_min_m = [] { const char* e = getenv("VIDCHAP_GEMM_PAIR_MIN_M"); return e ? atoi(e) : 256; }();
  if (pair_mode() && a->tile_n >= 0 && a->M >= pair_min_m) {
    const int mb2 = (a->M + 255) / 256;
    int BN2 = 0;
    for (int cand = 256; cand >= 128; cand >>= 1) {
      const long long tiles = (long long)mb2 * ((a->N + cand - 1) / cand) * p.splits;
      if (tiles >= (num_sms() / 2) || (cand == 128 && tiles >= num_sms() / 4)) { BN2 = cand; break; }
    }
    if (a->tile_n == 128 || a->tile_n == 256) BN2 = a->tile_n;
    if (BN2) {
      GemmParams p2 = p;
      p2.n_blocks = (a->N + BN2 - 1) / BN2;
      p2.ce_slots = 2 * p2.n_blocks;
      return launch_gemm_pair(a, BN2, em, p2, st);
    }
  }

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
    if (!a->a_mn_major && !a->b_mn_major) return launch_gemm<BN_, false, false>(tmA, tmB, em, p, grid, st); \
    if (!a->a_mn_major && a->b_mn_major) return launch_gemm<BN_, false, true>(tmA, tmB, em, p, grid, st);   \
    if (a->a_mn_major && !a->b_mn_major) return launch_gemm<BN_, true, false>(tmA, tmB, em, p, grid, st);   \
    return launch_gemm<BN_, true, true>(tmA, tmB, em, p, grid, st);                                        \
  }
  VC_DISPATCH(256)
  VC_DISPATCH(128)
  VC_DISPATCH(64)
#undef VC_DISPATCH
  return VC_ERR_INVALID;
}
