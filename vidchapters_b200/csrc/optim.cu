// Optimiser tail of the train step as flat-buffer HBM-bound kernels: global grad-norm, clip + Adam + bf16 weight
// re-pack in one pass, and the time-token embedding renorm.
//
// Reference: dvc.py:112-126 — torch.nn.utils.clip_grad_norm_(params, max_norm) (global L2, coef = min(1, max/(norm+1e-6))),
// torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) (dvc.py:345-351), then
//   W[-num_bins:] /= mean||W[-num_bins:]|| / mean||W[:-num_bins]||   for shared and lm_head (the same tensor when tied).
// All parameters live in ONE fp32 buffer (with m, v, grad buffers of the same layout and a bf16 shadow used by the
// GEMMs), so the whole tail is three streaming passes: ~ (4+4+8 read, 4+8+2 write) bytes per parameter.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

// Deterministic grid reduction: every block writes its partial sums, the last block to finish (atomic ticket) adds
// them up in index order.  The result must be bit-identical on every data-parallel rank: the clip coefficient and the
// time-token renorm factor derive from these sums, and ranks that round them differently drift apart (float atomics in
// arrival order did exactly that: tools/dp_check_gpu.py).  One reduction of each kind in flight at a time (one stream).
constexpr int kMaxRedBlocks = 4096;
__device__ float g_red_part[2][kMaxRedBlocks];
__device__ unsigned int g_red_ticket[2];

// vals[NV] are this block's partials (valid in thread 0); out[k] += total_k.  `slot` selects the scratch (0 sumsq, 1 rownorm).
template <int NV>
__device__ __forceinline__ void grid_reduce_ordered(const float (&vals)[NV], int slot, float* out) {
  __shared__ bool last;
  float* part = &g_red_part[0][0] + slot * kMaxRedBlocks;   // NV * gridDim.x <= kMaxRedBlocks per slot
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) part[k * gridDim.x + blockIdx.x] = vals[k];
    __threadfence();
    last = atomicInc(&g_red_ticket[slot], gridDim.x - 1) == gridDim.x - 1;   // wraps to 0: ready for the next launch
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float t = 0.f;
      for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) t += __ldcg(part + k * gridDim.x + i);   // fixed order per lane
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);                        // fixed tree
      if (threadIdx.x == 0) out[k] += t;
    }
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) s += g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t[1] = {0.f};
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) t[0] += red[i];
  grid_reduce_ordered<1>(t, 0, out);
}

// p,m,v updated in place; g is multiplied by grad_scale*clip_coef on the fly.  bias corrections follow torch.optim.Adam:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __nv_bfloat16* __restrict__ p_bf16, long long n, float lr, float b1, float b2, float eps, float bc1,
            float rsqrt_bc2, const float* __restrict__ norm_sq, float clip_max_norm, float grad_scale) {
  pdl_wait();
  pdl_trigger();
  float coef = grad_scale;
  if (norm_sq && clip_max_norm > 0.f) {
    const float total = sqrtf(*norm_sq) * grad_scale;
    coef *= fminf(1.0f, clip_max_norm / (total + 1e-6f));
  }
  const float step = lr / bc1;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * coef;
      mp[j] = b1 * mp[j] + (1.0f - b1) * gj;
      vp[j] = b2 * vp[j] + (1.0f - b2) * gj * gj;
      pp[j] -= step * mp[j] / (sqrtf(vp[j]) * rsqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p_bf16) reinterpret_cast<uint2*>(p_bf16)[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
}

// sums[0] += sum of row norms over rows [0, V-nb);  sums[1] += over rows [V-nb, V).  One warp per row.
__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ w, int V, int d, int nb, float* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int wt = gridDim.x * (blockDim.x >> 5);
  float acc0 = 0.f, acc1 = 0.f;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < V; r += wt) {
    const float4* row = reinterpret_cast<const float4*>(w + (long long)r * d);
    float s = 0.f;
    for (int c = lane; c < d / 4; c += 32) { const float4 x = row[c]; s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = sqrtf(s);
    if (r < V - nb) acc0 += nrm; else acc1 += nrm;
  }
  __shared__ float red[2][8];
  if (lane == 0) { red[0][threadIdx.x >> 5] = acc0; red[1][threadIdx.x >> 5] = acc1; }
  __syncthreads();
  float t[2] = {0.f, 0.f};
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) { t[0] += red[0][i]; t[1] += red[1][i]; }
  grid_reduce_ordered<2>(t, 1, sums);
}
// rows [V-nb, V) /= (mean_train / mean_frozen)
__global__ void renorm_apply_kernel(float* __restrict__ w, __nv_bfloat16* __restrict__ w_bf16, int V, int d, int nb,
                                    const float* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const float frozen = sums[0] / (float)(V - nb);
  const float train = sums[1] / (float)nb;
  const float div = train / frozen;
  const long long base = (long long)(V - nb) * d;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb * d) {
    const float x = w[base + i] / div;
    w[base + i] = x;
    if (w_bf16) w_bf16[base + i] = __float2bfloat16(x);
  }
}

__global__ void cast_flat_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) dst[i] = __float2bfloat16(src[i]);
}

}  // namespace vc

using namespace vc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int vc_sumsq(const float* g, int64_t n, float* out_accum, void* stream) {
  VC_CHECK(n > 0 && ((uintptr_t)g & 15) == 0, "vc_sumsq: bad args");
  VC_CUDA(launch_kernel(sumsq_kernel, dim3(num_sms() * 8), dim3(256), 0, ST(stream), g, n, out_accum));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                            float beta2, float eps, int step, const float* norm_sq, float clip_max_norm, float grad_scale,
                            void* stream) {
  VC_CHECK(n > 0 && n % 4 == 0 && step >= 1, "vc_adam_step: n must be a multiple of 4, step >= 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  VC_CUDA(launch_kernel(adam_kernel, dim3(num_sms() * 8), dim3(256), 0, ST(stream), p, g, m, v, (__nv_bfloat16*)p_bf16, n, lr, beta1, beta2, eps, (float)bc1,
                                                      (float)(1.0 / sqrt(bc2)), norm_sq, clip_max_norm, grad_scale));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_renorm_time_tokens(float* w, void* w_bf16, int V, int d, int num_bins, float* scratch2, void* stream) {
  VC_CHECK(V > num_bins && num_bins > 0 && d % 4 == 0, "vc_renorm_time_tokens: bad dims");
  VC_CUDA(cudaMemsetAsync(scratch2, 0, 2 * sizeof(float), ST(stream)));
  VC_CUDA(launch_kernel(rownorm_kernel, dim3(num_sms() * 4), dim3(256), 0, ST(stream), w, V, d, num_bins, scratch2));
  VC_CUDA(launch_kernel(renorm_apply_kernel, dim3((num_bins * d + 255) / 256), dim3(256), 0, ST(stream), w, (__nv_bfloat16*)w_bf16, V, d, num_bins, scratch2));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_cast_flat_bf16(const float* src, void* dst, int64_t n, void* stream) {
  VC_CHECK(n > 0, "vc_cast_flat_bf16: n");
  VC_CUDA(launch_kernel(cast_flat_kernel, dim3(num_sms() * 8), dim3(256), 0, ST(stream), src, (__nv_bfloat16*)dst, n));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
