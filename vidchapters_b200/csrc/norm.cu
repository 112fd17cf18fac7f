// Row-wise normalisation kernels (HBM-bound): T5 RMS LayerNorm and nn.LayerNorm, forward and backward.
//
// Reference: model/modeling_t5.py:254-277 (T5LayerNorm: x * rsqrt(mean(x^2) + eps) * w, fp32 statistics, no mean
// subtraction, no bias) and torch.nn.LayerNorm as used by model/vit.py:64,70,96 (eps 1e-5, affine).
// Residual stream x is fp32 in HBM; the normalised output feeds a tensor-core GEMM and is therefore written as bf16
// (optionally into a strided/batched destination so the encoder's final norm lands directly inside the decoder's
// concatenated [video ; text] memory, model/vid2seq.py:78).
//
// One warp per row, float4 loads (row = D/128 float4 per lane), warp-shuffle reductions; grid = enough CTAs to cover
// the SMs several times, grid-stride over rows.  Backward accumulates dw (and db) per warp in registers across its
// rows, reduces across the CTA's warps in shared memory, then one atomicAdd per column per CTA.
#include <cuda_bf16.h>

#include <cstdlib>
#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace vc {

constexpr int kMaxV4 = 8;  // D <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct RowMap {  // destination row of logical row r: (r / rows_per_batch) * batch_stride + row_offset + r % rows_per_batch
  int rows_per_batch, batch_stride, row_offset;
  __device__ __forceinline__ long long operator()(int r) const {
    return rows_per_batch > 0 ? (long long)(r / rows_per_batch) * batch_stride + row_offset + (r % rows_per_batch) : r;
  }
};

template <bool LAYERNORM>
__global__ void __launch_bounds__(256)
norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ out, float* __restrict__ out_f32, float* __restrict__ rstd_out,
                float* __restrict__ mean_out, int M, int D, float eps, float out_scale, RowMap map, uint32_t drop_seed_in,
                uint32_t drop_p16, const uint32_t* salt) {
  pdl_wait();
  pdl_trigger();
  const uint32_t drop_seed = drop_salted(drop_seed_in, salt);
  const int lane = threadIdx.x & 31;
  const int nv = D / 128;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps_total) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    float4 v[kMaxV4];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
      if (i < nv) {
        v[i] = xr[lane + 32 * i];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    float mean = 0.f, rstd;
    if (LAYERNORM) {
      mean = warp_sum(s) / D;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxV4; ++i) {
        if (i < nv) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
          var += a * a + b * b + c * c + d * d;
        }
      }
      rstd = rsqrtf(warp_sum(var) / D + eps);
    } else {
      rstd = rsqrtf(warp_sum(ss) / D + eps);
    }
    if (lane == 0) {
      if (rstd_out) rstd_out[row] = rstd;
      if (LAYERNORM && mean_out) mean_out[row] = mean;
    }
    const long long orow = map(row);
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        const float4 wv = *reinterpret_cast<const float4*>(w + c);
        float4 y;
        y.x = (v[i].x - mean) * rstd * wv.x; y.y = (v[i].y - mean) * rstd * wv.y;
        y.z = (v[i].z - mean) * rstd * wv.z; y.w = (v[i].w - mean) * rstd * wv.w;
        if (LAYERNORM) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + c);
          y.x += bv.x; y.y += bv.y; y.z += bv.z; y.w += bv.w;
        }
        y.x *= out_scale; y.y *= out_scale; y.z *= out_scale; y.w *= out_scale;
        if (drop_p16) {  // dropout on the normalised output (T5Stack final dropout, modeling_t5.py:1114)
          drop_apply<4>(&y.x, drop_row_key(drop_seed, (unsigned long long)row), drop_p16, (uint32_t)c, drop_scale(drop_p16));
        }
        if (out)
          *reinterpret_cast<uint2*>(out + orow * D + c) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + orow * D + c) = y;
      }
    }
  }
}

// Backward.  g = dL/dy (fp32, read through the same RowMap as the forward output), y = ((x-mean)*rstd*w + b)*scale.
//   dxhat = g*w*scale;  dx = rstd*(dxhat - mean(dxhat) [LN only] - xhat*mean(dxhat*xhat));  dx_accum (+)= dx
//   dw += sum_rows g*xhat*scale;  db += sum_rows g*scale
template <bool LAYERNORM, bool G_BF16>
__global__ void __launch_bounds__(256)
norm_bwd_kernel(const void* __restrict__ g_raw, const float* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ rstd_in, const float* __restrict__ mean_in, float* __restrict__ dx,
                __nv_bfloat16* __restrict__ dx_bf16, int accumulate_dx, float* __restrict__ dw, float* __restrict__ db,
                int M, int D, float scale, RowMap map, uint32_t g_drop_seed_in, uint32_t g_drop_p16, uint32_t dxb_drop_seed_in,
                uint32_t dxb_drop_p16, const uint32_t* salt) {
  pdl_wait();
  pdl_trigger();
  const uint32_t g_drop_seed = drop_salted(g_drop_seed_in, salt), dxb_drop_seed = drop_salted(dxb_drop_seed_in, salt);
  __shared__ float red[8][kMaxV4 * 128 + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = D / 128;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  float4 dw_acc[kMaxV4], db_acc[kMaxV4];
#pragma unroll
  for (int i = 0; i < kMaxV4; ++i) { dw_acc[i] = make_float4(0, 0, 0, 0); db_acc[i] = make_float4(0, 0, 0, 0); }
  for (int row = blockIdx.x * (blockDim.x >> 5) + warp; row < M; row += warps_total) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    const long long grow = map(row);
    const float rstd = rstd_in[row];
    const float mean = LAYERNORM ? mean_in[row] : 0.f;
    float4 xh[kMaxV4], dh[kMaxV4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
      if (i < nv) {
        const float4 xv = xr[lane + 32 * i];
        float4 gv;
        if (G_BF16) {
          const uint2 gb = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g_raw) + grow * D)[lane + 32 * i];
          gv = make_float4(bf16_lo(gb.x), bf16_hi(gb.x), bf16_lo(gb.y), bf16_hi(gb.y));
        } else {
          gv = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g_raw) + grow * D)[lane + 32 * i];
        }
        if (g_drop_p16) {  // the forward dropped the normalised output: the incoming gradient passes the same mask
          drop_apply<4>(&gv.x, drop_row_key(g_drop_seed, (unsigned long long)row), g_drop_p16, (uint32_t)((lane + 32 * i) * 4),
                        drop_scale(g_drop_p16));
        }
        const float4 wv = *reinterpret_cast<const float4*>(w + (lane + 32 * i) * 4);
        xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd; xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
        const float gx = gv.x * scale, gy = gv.y * scale, gz = gv.z * scale, gw = gv.w * scale;
        dh[i].x = gx * wv.x; dh[i].y = gy * wv.y; dh[i].z = gz * wv.z; dh[i].w = gw * wv.w;
        dw_acc[i].x += gx * xh[i].x; dw_acc[i].y += gy * xh[i].y; dw_acc[i].z += gz * xh[i].z; dw_acc[i].w += gw * xh[i].w;
        if (LAYERNORM) { db_acc[i].x += gx; db_acc[i].y += gy; db_acc[i].z += gz; db_acc[i].w += gw; }
        s1 += dh[i].x + dh[i].y + dh[i].z + dh[i].w;
        s2 += dh[i].x * xh[i].x + dh[i].y * xh[i].y + dh[i].z * xh[i].z + dh[i].w * xh[i].w;
      }
    }
    const float m1 = LAYERNORM ? warp_sum(s1) / D : 0.f;
    const float m2 = warp_sum(s2) / D;
    float4* dxr = reinterpret_cast<float4*>(dx + (long long)row * D);
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
      if (i < nv) {
        float4 o;
        o.x = rstd * (dh[i].x - m1 - xh[i].x * m2); o.y = rstd * (dh[i].y - m1 - xh[i].y * m2);
        o.z = rstd * (dh[i].z - m1 - xh[i].z * m2); o.w = rstd * (dh[i].w - m1 - xh[i].w * m2);
        if (accumulate_dx) {
          const float4 old = dxr[lane + 32 * i];
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        dxr[lane + 32 * i] = o;
        if (dx_bf16) {
          // bf16 copy = the dY of the sub-layer below, whose output went through dropout before the residual add
          if (dxb_drop_p16) {
            drop_apply<4>(&o.x, drop_row_key(dxb_drop_seed, (unsigned long long)row), dxb_drop_p16,
                          (uint32_t)((lane + 32 * i) * 4), drop_scale(dxb_drop_p16));
          }
          *reinterpret_cast<uint2*>(dx_bf16 + (long long)row * D + (lane + 32 * i) * 4) =
              make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
      }
    }
  }
  // cross-warp reduction of dw / db, then one atomic per column per CTA
  for (int pass = 0; pass < (LAYERNORM ? 2 : 1); ++pass) {
    if (pass == 1 && !db) break;
    if (pass == 0 && !dw) continue;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
      if (i < nv) {
        const float4 a = pass == 0 ? dw_acc[i] : db_acc[i];
        float* dst = &red[warp][(lane + 32 * i) * 4];
        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
      }
    }
    __syncthreads();
    float* target = pass == 0 ? dw : db;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += red[wv][c];
      atomicAdd(target + c, t);
    }
  }
}

// Backward for the model widths (D = 128*NV, NV = 6 or 8).  Same arithmetic as norm_bwd_kernel; laid out for HBM:
// every global load of a row (x, g and, when accumulating, the old dx) is issued before the first use, so one warp has
// ~7.7 KB in flight per row instead of three dependent load phases; the per-warp dw/db partial sums live in the warp's
// own shared-memory slice (no bank conflicts: lane l owns float4 slots l, l+32, ...) so the row data fits in registers
// without spills at 2 CTAs (16 warps) per SM — measured faster than 3 CTAs/SM with spills (37 vs 41 us at M=16000, D=768;
// the one-load-phase-at-a-time kernel it replaces took 70 us).
template <int NV, bool LAYERNORM, bool G_BF16>
__global__ void __launch_bounds__(256, 2)
norm_bwd_wide_kernel(const void* __restrict__ g_raw, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ rstd_in, const float* __restrict__ mean_in, float* __restrict__ dx,
                     __nv_bfloat16* __restrict__ dx_bf16, int accumulate_dx, float* __restrict__ dw, float* __restrict__ db,
                     int M, float scale, RowMap map, uint32_t g_drop_seed_in, uint32_t g_drop_p16, uint32_t dxb_drop_seed_in,
                     uint32_t dxb_drop_p16, const uint32_t* salt) {
  constexpr int D = NV * 128;
  extern __shared__ float4 wide_acc[];  // [LAYERNORM ? 2 : 1][8 warps][NV * 32]
  pdl_wait();
  pdl_trigger();
  const uint32_t g_drop_seed = drop_salted(g_drop_seed_in, salt), dxb_drop_seed = drop_salted(dxb_drop_seed_in, salt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* dw_acc = wide_acc + warp * (NV * 32);
  float4* db_acc = wide_acc + (8 + warp) * (NV * 32);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dw_acc[lane + 32 * i] = make_float4(0, 0, 0, 0);
    if (LAYERNORM) db_acc[lane + 32 * i] = make_float4(0, 0, 0, 0);
  }
  const int warps_total = gridDim.x * 8;
  using GVec = typename std::conditional<G_BF16, uint2, float4>::type;
  for (int row = blockIdx.x * 8 + warp; row < M; row += warps_total) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    float4* dxr = reinterpret_cast<float4*>(dx + (long long)row * D);
    const GVec* gr = reinterpret_cast<const GVec*>(reinterpret_cast<const char*>(g_raw) +
                                                   map(row) * (long long)D * (G_BF16 ? 2 : 4));
    float4 xh[NV], od[NV];
    GVec gq[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) xh[i] = xr[lane + 32 * i];
#pragma unroll
    for (int i = 0; i < NV; ++i) gq[i] = gr[lane + 32 * i];
    if (accumulate_dx) {
#pragma unroll
      for (int i = 0; i < NV; ++i) od[i] = dxr[lane + 32 * i];
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) od[i] = make_float4(0, 0, 0, 0);
    }
    const float rstd = rstd_in[row];
    const float mean = LAYERNORM ? mean_in[row] : 0.f;
    // od becomes old_dx + rstd * dxhat here, so the second pass needs only xh and od: dx = od - rstd*(m1 + xh*m2)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 gv, dh;
      if constexpr (G_BF16) gv = make_float4(bf16_lo(gq[i].x), bf16_hi(gq[i].x), bf16_lo(gq[i].y), bf16_hi(gq[i].y));
      else gv = gq[i];
      if (g_drop_p16) {
        drop_apply<4>(&gv.x, drop_row_key(g_drop_seed, (unsigned long long)row), g_drop_p16, (uint32_t)((lane + 32 * i) * 4),
                      drop_scale(g_drop_p16));
      }
      const float4 wv = *reinterpret_cast<const float4*>(w + (lane + 32 * i) * 4);
      xh[i].x = (xh[i].x - mean) * rstd; xh[i].y = (xh[i].y - mean) * rstd;
      xh[i].z = (xh[i].z - mean) * rstd; xh[i].w = (xh[i].w - mean) * rstd;
      const float gx = gv.x * scale, gy = gv.y * scale, gz = gv.z * scale, gw = gv.w * scale;
      dh.x = gx * wv.x; dh.y = gy * wv.y; dh.z = gz * wv.z; dh.w = gw * wv.w;
      od[i].x += rstd * dh.x; od[i].y += rstd * dh.y; od[i].z += rstd * dh.z; od[i].w += rstd * dh.w;
      float4 a = dw_acc[lane + 32 * i];
      a.x += gx * xh[i].x; a.y += gy * xh[i].y; a.z += gz * xh[i].z; a.w += gw * xh[i].w;
      dw_acc[lane + 32 * i] = a;
      if (LAYERNORM) {
        float4 b = db_acc[lane + 32 * i];
        b.x += gx; b.y += gy; b.z += gz; b.w += gw;
        db_acc[lane + 32 * i] = b;
      }
      s1 += dh.x + dh.y + dh.z + dh.w;
      s2 += dh.x * xh[i].x + dh.y * xh[i].y + dh.z * xh[i].z + dh.w * xh[i].w;
      asm volatile("" ::: "memory");  // keep the w / accumulator loads of chunk i+1 from being hoisted (register budget)
    }
    const float m1 = LAYERNORM ? rstd * warp_sum(s1) / D : 0.f;
    const float m2 = rstd * warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = od[i].x - m1 - xh[i].x * m2; o.y = od[i].y - m1 - xh[i].y * m2;
      o.z = od[i].z - m1 - xh[i].z * m2; o.w = od[i].w - m1 - xh[i].w * m2;
      dxr[lane + 32 * i] = o;
      if (dx_bf16) {
        if (dxb_drop_p16) {
          drop_apply<4>(&o.x, drop_row_key(dxb_drop_seed, (unsigned long long)row), dxb_drop_p16,
                        (uint32_t)((lane + 32 * i) * 4), drop_scale(dxb_drop_p16));
        }
        *reinterpret_cast<uint2*>(dx_bf16 + (long long)row * D + (lane + 32 * i) * 4) =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
  __syncthreads();
  const float* accf = reinterpret_cast<const float*>(wide_acc);
  for (int pass = 0; pass < (LAYERNORM ? 2 : 1); ++pass) {
    float* target = pass == 0 ? dw : db;
    if (!target) continue;
    for (int c = threadIdx.x; c < D; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += accf[(pass * 8 + wv) * D + c];
      atomicAdd(target + c, t);
    }
  }
}

template <int NV, bool LN, bool GB>
static cudaError_t launch_norm_bwd_wide(cudaStream_t st, const void* g, const float* x, const float* w, const float* rstd,
                                        const float* mean, float* dx, __nv_bfloat16* dx_bf16, int accumulate_dx, float* dw,
                                        float* db, int M, float scale, RowMap map, uint32_t s0, uint32_t p0, uint32_t s1,
                                        uint32_t p1) {
  auto kern = norm_bwd_wide_kernel<NV, LN, GB>;
  const int smem = (LN ? 2 : 1) * 8 * NV * 32 * (int)sizeof(float4);
  static PerDeviceOnce attr_done;  // per instantiation
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  // whole rows per warp, balanced: every warp of the grid gets the same number of rows (+-1 on the last CTA)
  const int blocks = (M + 7) / 8, cap = num_sms() * 2;
  const int rows_per_warp = (blocks + cap - 1) / cap;
  const int grid = (blocks + rows_per_warp - 1) / rows_per_warp;
  return launch_kernel(kern, dim3(grid), dim3(256), smem, st, g, x, w, rstd, mean, dx, dx_bf16, accumulate_dx, dw, db, M, scale,
                       map, s0, p0, s1, p1, drop_salt_ptr());
}

static int norm_grid(int M) {
  const int blocks = (M + 7) / 8;
  const int cap = num_sms() * 4;
  return blocks < cap ? blocks : cap;
}

}  // namespace vc

using namespace vc;

extern "C" int vc_norm_fwd(int kind, const float* x, const float* w, const float* bias, void* out_bf16, float* out_f32,
                           float* rstd, float* mean, int M, int D, float eps, float out_scale, int rows_per_batch,
                           int out_batch_stride, int out_row_offset, uint32_t drop_seed, uint32_t drop_p16, void* stream) {
  VC_CHECK(M > 0 && D > 0 && D % 128 == 0 && D <= 1024, "vc_norm_fwd: D=%d must be a multiple of 128 and <= 1024", D);
  VC_CHECK(kind == 0 || kind == 1, "vc_norm_fwd: kind 0=rms 1=layernorm");
  VC_CHECK(kind == 0 || bias != nullptr, "vc_norm_fwd: layernorm needs bias");
  RowMap map{rows_per_batch, out_batch_stride, out_row_offset};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (kind == 0)
    VC_CUDA(launch_kernel(norm_fwd_kernel<false>, dim3(norm_grid(M)), dim3(256), 0, st, x, w, bias, (__nv_bfloat16*)out_bf16, out_f32, rstd, mean, M, D,
                                                         eps, out_scale, map, drop_seed, drop_p16, drop_salt_ptr()));
  else
    VC_CUDA(launch_kernel(norm_fwd_kernel<true>, dim3(norm_grid(M)), dim3(256), 0, st, x, w, bias, (__nv_bfloat16*)out_bf16, out_f32, rstd, mean, M, D,
                                                        eps, out_scale, map, drop_seed, drop_p16, drop_salt_ptr()));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

extern "C" int vc_norm_bwd(int kind, const void* g, int g_bf16, const float* x, const float* w, const float* rstd, const float* mean,
                           float* dx, void* dx_bf16, int accumulate_dx, float* dw, float* db, int M, int D, float scale,
                           int rows_per_batch, int g_batch_stride, int g_row_offset, uint32_t g_drop_seed, uint32_t g_drop_p16,
                           uint32_t dxb_drop_seed, uint32_t dxb_drop_p16, void* stream) {
  VC_CHECK(M > 0 && D > 0 && D % 128 == 0 && D <= 1024, "vc_norm_bwd: D=%d must be a multiple of 128 and <= 1024", D);
  VC_CHECK(kind == 0 || kind == 1, "vc_norm_bwd: kind 0=rms 1=layernorm");
  RowMap map{rows_per_batch, g_batch_stride, g_row_offset};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool wide = [] { const char* e = getenv("VIDCHAP_NORM_BWD_WIDE"); return !(e && e[0] == '0'); }();
  if (wide && (D == 768 || D == 1024)) {
    __nv_bfloat16* dxb = (__nv_bfloat16*)dx_bf16;
#define VC_NBW(NV, LN, GB)                                                                                              \
  VC_CUDA((launch_norm_bwd_wide<NV, LN, GB>(st, g, x, w, rstd, mean, dx, dxb, accumulate_dx, dw, db, M, scale, map,     \
                                            g_drop_seed, g_drop_p16, dxb_drop_seed, dxb_drop_p16)))
    if (D == 768) {
      if (kind == 0) { if (g_bf16) VC_NBW(6, false, true); else VC_NBW(6, false, false); }
      else           { if (g_bf16) VC_NBW(6, true, true); else VC_NBW(6, true, false); }
    } else {
      if (kind == 0) { if (g_bf16) VC_NBW(8, false, true); else VC_NBW(8, false, false); }
      else           { if (g_bf16) VC_NBW(8, true, true); else VC_NBW(8, true, false); }
    }
#undef VC_NBW
    VC_CUDA(cudaGetLastError());
    return VC_OK;
  }
#define VC_NB(LN, GB)                                                                                                   \
  VC_CUDA(launch_kernel(norm_bwd_kernel<LN, GB>, dim3(norm_grid(M)), dim3(256), 0, st, g, x, w, rstd, mean, dx,         \
                        (__nv_bfloat16*)dx_bf16, accumulate_dx, dw, db, M, D, scale, map, g_drop_seed, g_drop_p16,      \
                        dxb_drop_seed, dxb_drop_p16, drop_salt_ptr()))
  if (kind == 0) { if (g_bf16) VC_NB(false, true); else VC_NB(false, false); }
  else           { if (g_bf16) VC_NB(true, true); else VC_NB(true, false); }
#undef VC_NB
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
