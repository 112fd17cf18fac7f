// 2-CTA (cta_group::2) variant of the bf16 tcgen05 GEMM: a CTA pair (cluster of 2, same TPC) computes one 256 x BN tile.
//
// Why: with one CTA per 128 x 256 tile every k-block pulls 48 KB through L2->SM for 2 M MACs (96 B/clk/SM at full
// tensor rate) and the kernel is L2-bandwidth bound at ~60 % of the tensor peak (profiles/r01_gemm_shapes_*.txt).  In
// pair mode each CTA stages only ITS 128 rows of A and HALF of the B tile (32 KB per k-block for the same MACs per SM,
// 64 B/clk/SM); the UMMA (M = 256) reads the two B halves from both CTAs' shared memory.
//
// Protocol (rank 0 = leader):
//   warp 0 lane 0, both CTAs : TMA producer; all loads complete_tx on the LEADER's full barrier (peer bit masked);
//                              the leader arms it with the pair's byte count.
//   warp 1 lane 0, leader    : tcgen05.mma.cta_group::2 (256 x BN x 16); commits multicast to BOTH CTAs' empty
//                              barriers (frees both smem rings) and to both tmem_full barriers.
//   warp 2, both CTAs        : tcgen05.alloc.cta_group::2 (2 x BN columns, accumulator double buffer).
//   warps 4..11, both CTAs   : epilogue on the CTA's own 128 accumulator rows (gemm_common.cuh); arrive (remotely for
//                              the peer) on the leader's tmem_empty barrier.
#include "gemm_common.cuh"

namespace vc {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // non-.aligned forms: the role lanes of warps 0/1 reach the final barrier later than their sibling lanes
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// Arrive on `bar` of cluster CTA `cta`.  Default (.release.cta) semantics on purpose: the only thing the waiter (the MMA
// issuer) consumes is "these TMEM columns have been read", which tcgen05.wait::ld + tcgen05.fence::before_thread_sync
// order; a .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR per tile per warp (12 % of the epilogue warps'
// stall samples in profiles/r02_gemm_epilogue.md).
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

constexpr int kEpiWarps2 = 8;    // pair kernel: 2 warps per TMEM lane quadrant, each owns half of the tile's columns (16 warps with
                                 // one pipeline stage less measured slower: the mainloop needs the depth)
constexpr int kThreads2 = 128 + 32 * kEpiWarps2;

template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = 128 * BK * 2;        // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * BK * 2;   // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes;  // BN=256: 6, BN=128: 8
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps2 * kEpiStageBytes + 1024 + 512;
  static constexpr int kTmemCols = 2 * BN;
};

template <int BN, bool A_MN, bool B_MN, bool TMA_EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ EpiMaps em, const GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kABytes = Cfg::kABytes, kBBytes = Cfg::kBBytes;
  constexpr int BNH = BN / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* epi_stage = smem + kStages * Cfg::kStageBytes;   // one 4 KB staging tile per epilogue warp (TMA epilogue)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + kEpiWarps2 * kEpiStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* epi_bar = tmem_empty + 2;   // one per epilogue warp (residual tile landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + kEpiWarps2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);    // leader's: its own arrive.expect_tx; bytes of both CTAs' loads
      mbar_init(&empty_bar[i], 1);   // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * 32 * kEpiWarps2);  // leader's: the epilogue threads of both CTAs
    }
    for (int i = 0; i < kEpiWarps2; ++i) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  if (warp == 2) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();      // prologue above overlapped the previous kernel's tail; from here on global memory is read
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;

  // tile scheduler over 256-row blocks
  const int m_blocks2 = (p.M + 255) / 256;
  const int tiles_mn = m_blocks2 * p.n_blocks;
  const int num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int mn = t % tiles_mn, split = t / tiles_mn;
      const int m0 = (mn % m_blocks2) * 256 + (int)rank * 128;
      const int n0 = (mn / m_blocks2) * BN + (int)rank * BNH;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes));
        uint8_t* sa = smem_a + stage * kABytes;
        uint8_t* sb = smem_b + stage * kBBytes;
        const int k0 = kb * BK;
        if (!A_MN) {
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], k0, m0);  // box 64(k) x 128(m)
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sa + j * (BK * 128), &tmA, &full_bar[stage], m0 + j * 64, k0);
        }
        if (!B_MN) {
          tma_load_2d_2sm(sb, &tmB, &full_bar[stage], k0, n0);  // box 64(k) x BN/2(n)
        } else {
#pragma unroll
          for (int j = 0; j < BNH / 64; ++j) tma_load_2d_2sm(sb + j * (BK * 128), &tmB, &full_bar[stage], n0 + j * 64, k0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA) =====================
    constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int split = t / tiles_mn;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * kABytes), A_MN ? BK * 128 : 0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * kBBytes), B_MN ? BK * 128 : 0, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = adesc + (uint64_t)((A_MN ? k * 2048 : k * 32) >> 4);
          const uint64_t bd = bdesc + (uint64_t)((B_MN ? k * 2048 : k * 32) >> 4);
          tc_mma_bf16_2sm(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        tc_commit_2sm(&empty_bar[stage]);  // frees this stage in BOTH CTAs
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      tc_commit_2sm(&tmem_full[acc]);  // accumulator complete: wake both CTAs' epilogues
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;          // which share of the tile's columns
    constexpr int kChunks = BN / 32 / (kEpiWarps2 / 4);
    int acc = 0;
    uint32_t acc_phase = 0, epi_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int mn = t % tiles_mn;
      const int m0 = (mn % m_blocks2) * 256 + (int)rank * 128, n0 = (mn / m_blocks2) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
      if (p.act == ACT_CE_STATS) {
        // fused cross entropy, statistics pass: nothing is stored; this warp's half of the tile's columns is folded into
        // (max, sum exp, sum z) of its rows and written to the row's slot of this (column block, half)
        // (two slots per column block, as in the single-CTA kernel: parts 0 and 1 take a half each, 2 and 3 only arrive)
        const int half = part;
        float cm = -INFINITY, cs = 0.f, ct = 0.f;
#pragma unroll 1
        for (int c = 0; c < (part < 2 ? BN / 64 : 0); ++c) {
          const int cc = half * (BN / 64) + c;
          ce_stats_chunk(p, tmem_base + acc * BN + cc * 32 + ((uint32_t)(q * 32) << 16), row, row_ok, n0 + cc * 32, alpha, cm, cs, ct);
        }
        if (row_ok && part < 2) {
          float* dst = p.ce_stats + ((long long)row * p.ce_slots + (n0 / BN) * 2 + half) * 3;
          dst[0] = cm; dst[1] = cs; dst[2] = ct;
        }
      } else if (!(p.dbg & 1)) {
        const uint32_t tcol = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
        if constexpr (TMA_EPI) {
          uint8_t* stage = epi_stage + (warp - 4) * kEpiStageBytes;
#define VC_EPI_PRESET(I)                                                                                                \
  case I:                                                                                                               \
    _Pragma("unroll 1") for (int c = 0; c < kChunks; ++c) {                                                             \
      const int cc = part * kChunks + c;                                                                                \
      gemm_epilogue_chunk_tma<epi_preset_mask(I)>(p, em, tcol + cc * 32, m0 + q * 32, lane, n0 + cc * 32, alpha, stage, \
                                                  &epi_bar[warp - 4], epi_phase);                                       \
    }                                                                                                                   \
    break;
          switch (p.epi_preset) {
            VC_EPI_PRESET(0) VC_EPI_PRESET(1) VC_EPI_PRESET(2) VC_EPI_PRESET(3) VC_EPI_PRESET(4)
            default:
            VC_EPI_PRESET(5)
          }
#undef VC_EPI_PRESET
        } else {
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c) {
            const int cc = part * kChunks + c;
            gemm_epilogue_chunk(p, tcol + cc * 32, row, row_ok, n0 + cc * 32, alpha);
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cta(&tmem_empty[acc], 0);  // the leader's barrier (remote arrive from the peer CTA)
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMA_EPI && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs done with TMEM and with each other's shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, bool TMA_EPI>
static int launch_gemm2_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, const GemmParams& p, int clusters,
                        cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  auto kern = gemm2_bf16_kernel<BN, A_MN, B_MN, TMA_EPI>;
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  VC_CUDA(launch_kernel(kern, dim3(2 * clusters), dim3(kThreads2), Cfg::kSmemBytes, st, tmA, tmB, em, p));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, const GemmParams& p, int clusters,
                        cudaStream_t st) {
  return p.tma_epi ? launch_gemm2_t<BN, A_MN, B_MN, true>(tmA, tmB, em, p, clusters, st)
                   : launch_gemm2_t<BN, A_MN, B_MN, false>(tmA, tmB, em, p, clusters, st);
}

// Called by vc_gemm_bf16 (gemm.cu) when the pair kernel applies.  BN in {128, 256}.
int launch_gemm_pair(const vc_gemm_args* a, int BN, const EpiMaps& em, const GemmParams& p, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int s;
  if (!a->a_mn_major) s = make_tmap_2d(&tmA, a->A, a->K, a->M, a->lda, 64, 128);
  else                s = make_tmap_2d(&tmA, a->A, a->M, a->K, a->lda, 64, BK);
  if (s != VC_OK) return s;
  if (!a->b_mn_major) s = make_tmap_2d(&tmB, a->B, a->K, a->N, a->ldb, 64, BN / 2);
  else                s = make_tmap_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, BK);
  if (s != VC_OK) return s;
  const int m_blocks2 = (a->M + 255) / 256;
  const int num_tiles = m_blocks2 * p.n_blocks * p.splits;
  const int max_clusters = num_sms() / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
#define VC_DISPATCH2(BN_)                                                                               \
  if (BN == BN_) {                                                                                      \
    if (!a->a_mn_major && !a->b_mn_major) return launch_gemm2<BN_, false, false>(tmA, tmB, em, p, clusters, st); \
    if (!a->a_mn_major && a->b_mn_major) return launch_gemm2<BN_, false, true>(tmA, tmB, em, p, clusters, st);   \
    if (a->a_mn_major && !a->b_mn_major) return launch_gemm2<BN_, true, false>(tmA, tmB, em, p, clusters, st);   \
    return launch_gemm2<BN_, true, true>(tmA, tmB, em, p, clusters, st);                                    \
  }
  VC_DISPATCH2(256)
  VC_DISPATCH2(128)
#undef VC_DISPATCH2
  set_error("launch_gemm_pair: unsupported BN %d", BN);
  return VC_ERR_INVALID;
}

}  // namespace vc
