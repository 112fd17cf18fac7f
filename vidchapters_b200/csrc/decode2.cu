// Incremental-decoding hot ops (HBM / latency bound, M <= 128 rows): single-query attention over a KV cache and the
// skinny linear layers of one decode step.
//
// Why not the training kernels: at one query per sequence the tcgen05 attention kernel computes 128-row tiles of which
// 127 rows are padding and read the cross-attention K/V at 2.4 TB/s; the 128 x 64-tile GEMM runs N=768 outputs on 12
// CTAs and pays TMEM allocation, tensor-map fetches and a 12-deep dependent pipeline for 64 rows (13.7 us per launch,
// profiles/r02_decode_step_launches.txt).  A greedy step at batch 64 is 2.6 GB of K/V + 0.28 GB of weights:
// the kernels below just stream those bytes with enough CTAs in flight.
//
// Replaces modeling_t5.py:484-525,539-581 for `past_key_value` decoding (one new query), T5LayerNorm + nn.Linear (+ReLU,
// +residual) of modeling_t5.py:254-277,296-311,591-656 at M = batch rows.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

constexpr float kDLog2e = 1.4426950408889634f;
constexpr float kDMasked = -3.0e38f;

// ---------------------------------------------------------------------------------------------------------------------
// Single-query attention: one CTA = one (head, sequence).  256 threads.
//   phase 1: thread-per-key dot products q.k (K rows are 128-byte lines; 8 x 16-byte loads each), scores -> smem
//   phase 2: block max / sum (log2 domain, integer maximum like attn_fwd.cu so the bf16 rounding of P is the same)
//   phase 3: warp w accumulates keys w, w+8, ...: lane l owns output dims 2l, 2l+1 (one coalesced 128-byte V row per key)
// Same arithmetic and rounding points as attn_fwd.cu: p~ = 2^(s2 - ceil(max)) rounded to bf16 before it meets V, the
// row sum uses the unrounded p~, masked keys take the reference's additive finfo.min.
struct AttnDecParams {
  const __nv_bfloat16* q; long long ldq; int q_col;
  const __nv_bfloat16* k; long long ldk; int k_col;
  const __nv_bfloat16* v; long long ldv; int v_col;
  __nv_bfloat16* out; long long ldo;
  int B, H, Lk;
  long long kv_batch_rows;
  int kv_batch_div;
  const float* bias_rel; int bias_zero, bias_len;
  const uint8_t* kmask;
  int causal;
  float scale_log2e;
  int q_offset; const int* q_offset_dev;
};

__global__ void __launch_bounds__(256, 3) attn_decode_kernel(const AttnDecParams p) {
  extern __shared__ float s_sc[];            // [Lk] scores, then probabilities
  __shared__ float s_red[8];
  __shared__ float s_acc[8][64];
  pdl_wait();
  pdl_trigger();
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qpos = p.q_offset + (p.q_offset_dev ? __ldg(p.q_offset_dev) : 0);
  const int nk = p.causal ? min(p.Lk, qpos + 1) : p.Lk;          // causal: keys past the query do not exist yet
  const long long kvb = (long long)(b / p.kv_batch_div) * p.kv_batch_rows;
  // q (64 bf16 = 8 x uint4): every thread needs all of it — read as shared-memory broadcasts (held in 32 registers per
  // thread next to two keys' worth of loads, the loop spilled at the 80 registers that 3 CTAs per SM allow)
  __shared__ uint4 s_q[8];
  if (tid < 8) s_q[tid] = __ldg(reinterpret_cast<const uint4*>(p.q + (long long)b * p.ldq + p.q_col + h * 64) + tid);
  __syncthreads();
  const float* brow = p.bias_rel ? p.bias_rel + (long long)h * p.bias_len + (p.bias_zero - qpos) : nullptr;
  const uint8_t* mrow = p.kmask ? p.kmask + (long long)b * p.Lk : nullptr;
  float m_loc = -INFINITY;
#pragma unroll 2
  for (int k = tid; k < nk; k += 256) {
    const uint4* kp = reinterpret_cast<const uint4*>(p.k + (kvb + k) * p.ldk + p.k_col + h * 64);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 kk = __ldg(kp + i), qq = s_q[i];
      const uint32_t a[4] = {qq.x, qq.y, qq.z, qq.w}, c[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) acc = fmaf(bf16_lo(a[e]), bf16_lo(c[e]), fmaf(bf16_hi(a[e]), bf16_hi(c[e]), acc));
    }
    float s2 = acc * p.scale_log2e + (brow ? __ldg(brow + k) * kDLog2e : 0.f);
    if (mrow && mrow[k] == 0) s2 = kDMasked;
    s_sc[k] = s2;
    m_loc = fmaxf(m_loc, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m_loc = fmaxf(m_loc, __shfl_xor_sync(0xffffffffu, m_loc, o));
  if (lane == 0) s_red[warp] = m_loc;
  __syncthreads();
  float m = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
  m = ceilf(m);
  __syncthreads();
  float l_loc = 0.f;
  for (int k = tid; k < nk; k += 256) {
    const float pk = fast_exp2(s_sc[k] - m);
    l_loc += pk;
    s_sc[k] = __bfloat162float(__float2bfloat16(pk));     // the probability meets V rounded to bf16
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l_loc += __shfl_xor_sync(0xffffffffu, l_loc, o);
  if (lane == 0) s_red[warp] = l_loc;
  __syncthreads();
  float l = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) l += s_red[w];
  // phase 3: lane l = (key sub-index l / 8, dims 8 (l % 8) .. +8): a warp instruction reads FOUR 128-byte V rows; warp w
  // walks keys 4 (w + 8 i) + l / 8.  Unrolled x8: 32 rows = 4 KB per warp, 96 KB per SM in flight (the loop is pure HBM
  // latency; at x4 the V pass ran at about half the bandwidth of the K pass).
  float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int ksub = lane >> 3, dsub = (lane & 7) * 8;
  const __nv_bfloat16* vbase = p.v + kvb * p.ldv + p.v_col + h * 64 + dsub;
#pragma unroll 8
  for (int k = warp * 4 + ksub; k < nk; k += 32) {
    const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vbase + (long long)k * p.ldv));
    const float pk = s_sc[k];
    const uint32_t w4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o8[2 * e] = fmaf(pk, bf16_lo(w4[e]), o8[2 * e]);
      o8[2 * e + 1] = fmaf(pk, bf16_hi(w4[e]), o8[2 * e + 1]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {      // fold the four key sub-groups of the warp
    o8[e] += __shfl_xor_sync(0xffffffffu, o8[e], 8);
    o8[e] += __shfl_xor_sync(0xffffffffu, o8[e], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[warp][lane * 8 + e] = o8[e];
  }
  __syncthreads();
  if (tid < 64) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) o += s_acc[w][tid];
    p.out[(long long)b * p.ldo + h * 64 + tid] = __float2bfloat16(o / l);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Skinny linear layer: out[M, N] = epi( A[M, K] . W[N, K]^T ), M <= 64 rows per CTA row block (grid.y), 16 output columns
// per CTA (grid.x = N / 16), 4 warps = 4 x 16 rows, mma.sync m16n8k16 (bf16 in, fp32 accumulate).
//   A: bf16 [M][lda]  — or — the fp32 residual stream with the T5 RMS norm applied on the fly (x * rsqrt(mean x^2 + eps)
//      * nw * out_scale, rounded to bf16: exactly what vc_norm_fwd feeds the GEMM), so norm + linear is ONE launch.
//   epilogue: ReLU (bf16 out) | residual add into the fp32 stream (out may alias the residual) | plain bf16 / fp32 store.
// The whole K extent of A (K <= 1024) or a 1024-wide chunk of it sits in shared memory next to the CTA's 16 weight rows.
struct DecLinParams {
  const void* A; long long lda; int a_fp32;
  const float* norm_w; float eps, out_scale;
  const __nv_bfloat16* W; long long ldw;
  void* out; long long ldo; int out_fp32;
  const float* residual; long long ldr;
  int relu;
  int M, N, K;
};

constexpr int kDLKC = 1024;                 // K chunk held in shared memory
constexpr int kDLStride = kDLKC + 8;        // bf16 elements per smem row (+16 B: conflict-free ldmatrix)

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NT>   // n8 tiles per CTA: 2 (16 output columns) or 4 (32: when N / 16 CTAs would not fit the machine in one wave)
__global__ void __launch_bounds__(512) decode_linear_kernel(const DecLinParams p) {
  extern __shared__ __align__(16) uint8_t dl_smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(dl_smem);            // [64][kDLStride]
  __nv_bfloat16* sW = sA + 64 * kDLStride;                                  // [8 * NT][kDLStride]
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NC = 8 * NT;
  const int n0 = blockIdx.x * NC, m0 = blockIdx.y * 64;
  const int rows = min(64, p.M - m0);
  float acc[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  // The kernel is pure latency (a few MB per launch): 512 threads stage the operands with every load of a round issued
  // before the first use — ONE round for the 16 weight rows, one (bf16 input) or two (fp32 + norm) for the 64 input rows —
  // then warps 0..3 run the tensor-core loop.
  // gridDim.z > 1 (split-K, in-place residual layers with K > 1024): this CTA takes ONE 1024-wide chunk of K and adds its
  // partial sums onto the fp32 stream with atomics — 3x the CTAs for the K = 3072 projection that ran on 48 of 148 SMs.
  const bool split = gridDim.z > 1;
  const int kc_begin = split ? (int)blockIdx.z * kDLKC : 0;
  const int kc_end = split ? min(p.K, kc_begin + kDLKC) : p.K;
  for (int kc = kc_begin; kc < kc_end; kc += kDLKC) {
    const int kw = min(kDLKC, p.K - kc);
    if (kc > kc_begin) __syncthreads();
    {   // the CTA's weight rows go straight to shared memory (cp.async: no registers held across the input staging)
      const int nv = kw / 8;
      for (int i = tid; i < NC * nv; i += 512) {
        const int r = i / nv, c = i % nv;
        const bool ok = n0 + r < p.N;
        cp_async16(sW + r * kDLStride + c * 8, reinterpret_cast<const uint4*>(p.W + (long long)(ok ? n0 + r : 0) * p.ldw + kc) + c,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    }
    if (p.a_fp32) {
      // warp w owns rows 4w .. 4w+3 (two at a time): lane holds columns lane + 32 i of each row, so the RMS statistic is a
      // warp reduction and the normalised row goes to shared memory without a second pass over x
      const float* X = reinterpret_cast<const float*>(p.A);
      const int nv = kw / 4;       // <= 256 float4 per row -> 8 per lane
#pragma unroll 1
      for (int rr = 0; rr < 4; rr += 2) {
        float4 v[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = warp * 4 + rr + j, c = lane + 32 * i;
            v[j][i] = (r < rows && c < nv) ? __ldg(reinterpret_cast<const float4*>(X + (long long)(m0 + r) * p.lda + kc) + c)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = warp * 4 + rr + j;
          float rs = 1.0f;
          if (p.norm_w) {
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) ss += v[j][i].x * v[j][i].x + v[j][i].y * v[j][i].y + v[j][i].z * v[j][i].z + v[j][i].w * v[j][i].w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            rs = rsqrtf(ss / p.K + p.eps) * p.out_scale;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c >= nv) continue;
            float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.norm_w) w = __ldg(reinterpret_cast<const float4*>(p.norm_w + kc) + c);
            const float4 x = v[j][i];
            *reinterpret_cast<uint2*>(sA + r * kDLStride + c * 4) =
                make_uint2(pack_bf16x2(x.x * rs * w.x, x.y * rs * w.y), pack_bf16x2(x.z * rs * w.z, x.w * rs * w.w));
          }
        }
      }
    } else {
      const __nv_bfloat16* Ab = reinterpret_cast<const __nv_bfloat16*>(p.A);
      const int nv = kw / 8;       // <= 128 uint4 per row: 64 rows -> <= 8192 -> 16 per thread, one round
      uint4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = tid + 512 * j, r = i / nv, c = i % nv;
        v[j] = (i < 64 * nv && r < rows) ? __ldg(reinterpret_cast<const uint4*>(Ab + (long long)(m0 + r) * p.lda + kc) + c)
                                         : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = tid + 512 * j, r = i / nv, c = i % nv;
        if (i < 64 * nv) *reinterpret_cast<uint4*>(sA + r * kDLStride + c * 8) = v[j];
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (warp >= 4) continue;
    // ---- warp w: rows 16w..16w+15, all 16 columns: per k16 step one A fragment (x4) and both B fragments (x4)
    const uint32_t a_base = smem_u32(sA + (warp * 16 + (lane & 15)) * kDLStride + (lane >> 4) * 8);
    const uint32_t b_base = smem_u32(sW + ((lane & 7) + ((lane >> 4) << 3)) * kDLStride + ((lane >> 3) & 1) * 8);
    for (int k = 0; k < kw; k += 16) {
      uint32_t a0, a1, a2, a3;
      ldmatrix_x4(a0, a1, a2, a3, a_base + k * 2);
#pragma unroll
      for (int t = 0; t < NT; t += 2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, b_base + t * 8 * kDLStride * 2 + k * 2);   // (b0,b1): columns 8t..8t+7 ; (b2,b3): 8t+8..8t+15
        mma_bf16_16816(acc[t], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(acc[t + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  }
  // ---- epilogue (warps 0..3): thread holds rows (lane/4) and (lane/4 + 8) of its warp's 16, columns 2*(lane%4) + {0,1} of
  // each n8 tile
  if (warp >= 4) return;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int r = m0 + warp * 16 + (lane >> 2) + hlf * 8;
      const int c = n0 + t * 8 + (lane & 3) * 2;
      if (r >= p.M || c >= p.N) continue;
      float v0 = acc[t][hlf * 2], v1 = acc[t][hlf * 2 + 1];
      if (split) {   // out already holds the residual (out == residual)
        float* dst = reinterpret_cast<float*>(p.out) + (long long)r * p.ldo + c;
        atomicAdd(dst, v0);
        atomicAdd(dst + 1, v1);
        continue;
      }
      if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      if (p.residual) {
        const float2 rr = *reinterpret_cast<const float2*>(p.residual + (long long)r * p.ldr + c);
        v0 += rr.x; v1 += rr.y;
      }
      if (p.out_fp32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + (long long)r * p.ldo + c) = make_float2(v0, v1);
      else *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)r * p.ldo + c) = pack_bf16x2(v0, v1);
    }
  }
}

int launch_attn_decode(const vc_attn_args* a, cudaStream_t st) {
  AttnDecParams p;
  p.q = (const __nv_bfloat16*)a->q; p.ldq = a->ldq; p.q_col = a->q_col;
  p.k = (const __nv_bfloat16*)a->k; p.ldk = a->ldk; p.k_col = a->k_col;
  p.v = (const __nv_bfloat16*)a->v; p.ldv = a->ldv; p.v_col = a->v_col;
  p.out = (__nv_bfloat16*)a->out; p.ldo = a->ldo;
  p.B = a->B; p.H = a->H; p.Lk = a->Lk;
  p.kv_batch_rows = a->kv_batch_rows > 0 ? a->kv_batch_rows : a->Lk;
  p.kv_batch_div = a->kv_batch_div > 1 ? a->kv_batch_div : 1;
  p.bias_rel = a->bias_rel;
  p.bias_zero = a->bias_len > 0 ? a->bias_zero : 0;          // Lq = 1: training layout puts relative position 0 at index 0
  p.bias_len = a->bias_len > 0 ? a->bias_len : a->Lk;
  p.kmask = a->kmask; p.causal = a->causal;
  p.scale_log2e = a->scale * kDLog2e;
  p.q_offset = a->q_offset; p.q_offset_dev = a->q_offset_dev;
  VC_CHECK(a->Lk <= 12000, "vc_attn_fwd (decode): Lk=%d exceeds the score staging", a->Lk);
  static PerDeviceOnce attr;
  if (attr.need()) {
    VC_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48000));
  }
  VC_CUDA(launch_kernel(attn_decode_kernel, dim3(a->H, a->B), dim3(256), (size_t)a->Lk * 4, st, p));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc

using namespace vc;

extern "C" int vc_decode_linear(const void* A, int64_t lda, int a_fp32, const float* norm_w, float eps, float out_scale,
                                const void* W, int64_t ldw, void* out, int64_t ldo, int out_fp32, const float* residual,
                                int64_t ldr, int relu, int M, int N, int K, void* stream) {
  VC_CHECK(A && W && out && M > 0 && N > 0 && K > 0, "vc_decode_linear: bad arguments");
  VC_CHECK(K % 16 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 2 == 0 && ldr % 2 == 0, "vc_decode_linear: K x16, strides x8");
  VC_CHECK(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "vc_decode_linear: A/W must be 16-byte aligned");
  VC_CHECK(!norm_w || (a_fp32 && K <= kDLKC), "vc_decode_linear: the fused RMS norm needs fp32 input and K <= %d", kDLKC);
  VC_CHECK(K % 32 == 0, "vc_decode_linear: K must be a multiple of 32");
  VC_CHECK(!residual || out_fp32, "vc_decode_linear: residual add needs fp32 out");
  DecLinParams p;
  p.A = A; p.lda = lda; p.a_fp32 = a_fp32; p.norm_w = norm_w; p.eps = eps; p.out_scale = out_scale;
  p.W = (const __nv_bfloat16*)W; p.ldw = ldw; p.out = out; p.ldo = ldo; p.out_fp32 = out_fp32;
  p.residual = residual; p.ldr = ldr; p.relu = relu; p.M = M; p.N = N; p.K = K;
  // split-K only where the partial sums can land on the residual stream in place (fp32 atomics: the order of the <= 3 adds
  // per element is not fixed, i.e. the last bit of the stream may differ between runs)
  static const bool splitk_on = [] { const char* e = getenv("VIDCHAP_DECODE_SPLITK"); return !(e && e[0] == '0'); }();
  const bool inplace = residual && (const void*)residual == (const void*)out && ldr == ldo && out_fp32 && !relu && !norm_w;
  const int splits = (splitk_on && inplace && K > kDLKC) ? (K + kDLKC - 1) / kDLKC : 1;
  const int mblocks = (M + 63) / 64;
  // 32 columns per CTA when 16-column CTAs would need a second wave (N = 3072: 192 CTAs on 148 SMs, one CTA per SM)
  const bool wide = (long long)((N + 15) / 16) * mblocks * splits > num_sms();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (wide) {
    const size_t smem = (size_t)(64 + 32) * kDLStride * 2;
    static PerDeviceOnce attr;
    if (attr.need()) VC_CUDA(cudaFuncSetAttribute(decode_linear_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VC_CUDA(launch_kernel(decode_linear_kernel<4>, dim3((N + 31) / 32, mblocks, splits), dim3(512), smem, st, p));
  } else {
    const size_t smem = (size_t)(64 + 16) * kDLStride * 2;
    static PerDeviceOnce attr;
    if (attr.need()) VC_CUDA(cudaFuncSetAttribute(decode_linear_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VC_CUDA(launch_kernel(decode_linear_kernel<2>, dim3((N + 15) / 16, mblocks, splits), dim3(512), smem, st, p));
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
