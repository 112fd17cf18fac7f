// Small HBM-bound kernels of the Vid2Seq step: embedding gather / scatter-add, target preparation, relative-position
// bias expansion, positional-embedding add, label-smoothed cross-entropy (+ its gradient), column sums, casts.
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- embedding gather: out[i,:] = table[ids[i],:]   (model/vid2seq.py:71, modeling_t5.py:972)
__global__ void __launch_bounds__(256) embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                                       float* __restrict__ out, int n, int d, int V, uint32_t drop_seed_in,
                                                       uint32_t drop_p16, const uint32_t* salt) {
  pdl_wait();
  pdl_trigger();
  const uint32_t drop_seed = drop_salted(drop_seed_in, salt);
  const int lane = threadIdx.x & 31;
  const int wt = gridDim.x * (blockDim.x >> 5);
  const float sc = drop_p16 ? drop_scale(drop_p16) : 1.f;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += wt) {
    long long id = ids[i];
    // An id outside [0, V) is a caller bug: nn.Embedding raises a device assert for it (modeling_t5.py:972).  Here the row
    // is poisoned with NaN so that the loss turns non-finite and dvc.py:107-110 aborts the run — loud, with no extra
    // host synchronisation on the hot path.
    const bool bad = id < 0 || id >= V;
    if (bad) id = 0;
    const float4* src = reinterpret_cast<const float4*>(table + id * d);
    float4* dst = reinterpret_cast<float4*>(out + (long long)i * d);
    for (int c = lane; c < d / 4; c += 32) {
      float4 v = __ldg(src + c);
      if (bad) v = make_float4(NAN, NAN, NAN, NAN);
      if (drop_p16) {
        drop_apply<4>(&v.x, drop_row_key(drop_seed, (unsigned long long)i), drop_p16, (uint32_t)(c * 4), sc);
      }
      dst[c] = v;
    }
  }
}
// ---- embedding backward: dtable[ids[i],:] += dout[i,:]   (autograd of nn.Embedding; tied table)
__global__ void __launch_bounds__(256) embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dout,
                                                       float* __restrict__ dtable, int n, int d, int V, uint32_t drop_seed_in,
                                                       uint32_t drop_p16, const uint32_t* salt) {
  pdl_wait();
  pdl_trigger();
  const uint32_t drop_seed = drop_salted(drop_seed_in, salt);
  const int lane = threadIdx.x & 31;
  const int wt = gridDim.x * (blockDim.x >> 5);
  const float sc = drop_p16 ? drop_scale(drop_p16) : 1.f;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += wt) {
    long long id = ids[i];
    if (id < 0 || id >= V) continue;
    const float4* src = reinterpret_cast<const float4*>(dout + (long long)i * d);
    float4* dst = reinterpret_cast<float4*>(dtable + id * d);
    for (int c = lane; c < d / 4; c += 32) {
      float4 v = src[c];
      if (drop_p16) {
        drop_apply<4>(&v.x, drop_row_key(drop_seed, (unsigned long long)i), drop_p16, (uint32_t)(c * 4), sc);
      }
      atomicAdd(dst + c, v);
    }
  }
}

// ---- targets: labels = ids (pad -> -100) (vid2seq.py:86-88); dec_in = shift_right(labels) (modeling_t5.py:845-868);
//      n_valid = #labels != -100 (denominator of F.cross_entropy's mean, modeling_t5.py:1721)
__global__ void prepare_targets_kernel(const long long* __restrict__ out_ids, long long* __restrict__ dec_in,
                                       long long* __restrict__ labels, float* __restrict__ n_valid, int B, int S,
                                       long long pad_id) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int valid = 0;
  if (i < B * S) {
    const int s = i % S;
    const long long id = out_ids[i];
    labels[i] = (id == pad_id) ? -100 : id;
    valid = (id != pad_id);
    long long prev = 0;  // decoder_start_token_id = 0
    if (s > 0) {
      prev = out_ids[i - 1];
      if (prev == pad_id) prev = 0;  // -100 -> pad_token_id (=0)
    }
    dec_in[i] = prev;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_valid, (float)__popc(bal));
}

// ---- relative position bias: out[h][r] = table[lut[r]][h]   (modeling_t5.py:445-460; lut = bucket of r-(Lq-1))
__global__ void bias_expand_kernel(const float* __restrict__ table, const int* __restrict__ lut, float* __restrict__ out,
                                   int H, int R) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * R) {
    const int h = i / R, r = i % R;
    out[i] = table[lut[r] * H + h];
  }
}
// backward: dtable[lut[r]][h] += drel[h][r]
// One CTA per head.  T5's buckets send every distance beyond +-128 to ONE bucket, so a thread per element with a global
// atomicAdd serialised ~1700 adds on one address per head (120-180 us for 24 k elements in the graph timeline).  Here a
// thread keeps a running sum while its elements stay in one bucket, flushes into a shared-memory table, and the CTA issues
// one global atomic per (bucket, head).
constexpr int kFoldBuckets = 256;
__global__ void __launch_bounds__(256) bias_fold_kernel(const float* __restrict__ drel, const int* __restrict__ lut,
                                                        float* __restrict__ dtable, int H, int R) {
  __shared__ float tab[kFoldBuckets];
  pdl_wait();
  pdl_trigger();
  const int h = blockIdx.x, tid = threadIdx.x;
  tab[tid] = 0.f;
  __syncthreads();
  auto flush = [&](int b, float v) {
    if (v == 0.f) return;
    if (b < kFoldBuckets) atomicAdd(&tab[b], v);
    else atomicAdd(dtable + (long long)b * H + h, v);
  };
  int cur = -1;
  float acc = 0.f;
  for (int r = tid; r < R; r += 256) {
    const float g = drel[(long long)h * R + r];
    const int b = __ldg(lut + r);
    if (b != cur) {
      if (cur >= 0) flush(cur, acc);
      cur = b;
      acc = 0.f;
    }
    acc += g;
  }
  // the last run of most threads is the far bucket: when a whole warp ends in the same bucket, one lane adds the warp's sum
  const int cur0 = __shfl_sync(0xffffffffu, cur, 0);
  if (__all_sync(0xffffffffu, cur == cur0)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0 && cur >= 0) flush(cur, acc);
  } else if (cur >= 0) {
    flush(cur, acc);
  }
  __syncthreads();
  if (tab[tid] != 0.f) atomicAdd(dtable + (long long)tid * H + h, tab[tid]);
}

// ---- x + pos_embed (model/vit.py:119-127; nearest interpolation when T != num_features)
__global__ void add_pos_kernel(const float* __restrict__ x, const float* __restrict__ pos, float* __restrict__ out, int B,
                               int T, int C, int P, uint32_t drop_seed_in, uint32_t drop_p16, const uint32_t* salt) {
  pdl_wait();
  pdl_trigger();
  const uint32_t drop_seed = drop_salted(drop_seed_in, salt);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  const long long total = (long long)B * T * C / 4;
  if (i < total) {
    const int c4 = (int)(i % (C / 4));
    const int t = (int)((i / (C / 4)) % T);
    const int src_t = (T == P) ? t : (int)floorf((float)t * ((float)P / (float)T));
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos) + (long long)src_t * (C / 4) + c4);
    float4 v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    if (drop_p16) {
      // mask row = (b, t), column = channel
      drop_apply<4>(&v.x, drop_row_key(drop_seed, (unsigned long long)(i / (C / 4))), drop_p16, (uint32_t)(c4 * 4),
                    drop_scale(drop_p16));
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}
__global__ void add_pos_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dpos, int B, int T, int C, int P,
                                   uint32_t drop_seed_in, uint32_t drop_p16, const uint32_t* salt) {
  pdl_wait();
  pdl_trigger();
  const uint32_t drop_seed = drop_salted(drop_seed_in, salt);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over T*C
  if (i < T * C) {
    const int t = i / C, c = i % C;
    const int src_t = (T == P) ? t : (int)floorf((float)t * ((float)P / (float)T));
    float s = 0.f;
    const float sc = drop_p16 ? drop_scale(drop_p16) : 1.f;
    for (int b = 0; b < B; ++b) {
      const unsigned long long rowi = (unsigned long long)b * T + t;
      const float g = dx[rowi * C + c];
      s += (!drop_p16 || drop_keep(drop_row_key(drop_seed, rowi), drop_p16, (uint32_t)c)) ? g * sc : 0.f;
    }
    atomicAdd(dpos + src_t * C + c, s);
  }
}

// ---- label-smoothed cross entropy over fp32 logits (modeling_t5.py:1721: F.cross_entropy(ignore_index=-100,
//      label_smoothing=eps), mean over valid rows) and its gradient (w.r.t. logits, for upstream grad 1):
//      loss_row = (1-eps)*(lse - z_y) + eps*(lse - mean_c z_c);  dz_c = (softmax_c - (1-eps)[c==y] - eps/V) / n_valid
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                     const float* __restrict__ n_valid_p, float eps, float* __restrict__ loss_out,
                     __nv_bfloat16* __restrict__ dlogits, long long ldd, int V) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8];
  __shared__ float bcast[2];
  const int row = blockIdx.x;
  const long long y = labels[row];
  const float* z = logits + (long long)row * ld;
  __nv_bfloat16* dz = dlogits ? dlogits + (long long)row * ldd : nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V4 = V & ~3, V8 = V & ~7;
  if (y < 0) {  // ignore_index: zero gradient row
    if (dz) {
      for (int c = tid * 8; c < V8; c += 256 * 8) *reinterpret_cast<uint4*>(dz + c) = make_uint4(0, 0, 0, 0);
      for (int c = V8 + tid; c < V; c += 256) dz[c] = __float2bfloat16(0.f);
    }
    return;
  }
  float mx = -INFINITY;
  for (int c = tid * 4; c < V4; c += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(z + c);
    mx = fmaxf(fmaxf(fmaxf(mx, v.x), fmaxf(v.y, v.z)), v.w);
  }
  for (int c = V4 + tid; c < V; c += 256) mx = fmaxf(mx, z[c]);
  mx = warp_max_f(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) { float m = red[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]); bcast[0] = m; }
  __syncthreads();
  mx = bcast[0];
  float se = 0.f, sz = 0.f;
  for (int c = tid * 4; c < V4; c += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(z + c);
    se += __expf(v.x - mx) + __expf(v.y - mx) + __expf(v.z - mx) + __expf(v.w - mx);
    sz += v.x + v.y + v.z + v.w;
  }
  for (int c = V4 + tid; c < V; c += 256) { se += __expf(z[c] - mx); sz += z[c]; }
  se = warp_sum_f(se); sz = warp_sum_f(sz);
  __syncthreads();
  if (lane == 0) red[warp] = se;
  __syncthreads();
  if (tid == 0) { float s = 0; for (int i = 0; i < 8; ++i) s += red[i]; bcast[0] = s; }
  __syncthreads();
  se = bcast[0];
  __syncthreads();
  if (lane == 0) red[warp] = sz;
  __syncthreads();
  if (tid == 0) { float s = 0; for (int i = 0; i < 8; ++i) s += red[i]; bcast[1] = s; }
  __syncthreads();
  sz = bcast[1];
  const float lse = mx + logf(se);
  const float nv = *n_valid_p;
  if (tid == 0) {
    // a label >= V is a caller bug (F.cross_entropy raises a device assert): poison the loss instead of reading z[y]
    const float nll = y < V ? lse - z[y] : NAN;
    const float smooth = lse - sz / (float)V;
    atomicAdd(loss_out, ((1.0f - eps) * nll + eps * smooth) / nv);
  }
  if (dz) {
    const float inv_nv = 1.0f / nv, inv_se = 1.0f / se, sm = eps / (float)V;
    for (int c = tid * 8; c < V8; c += 256 * 8) {
      const float4 a = *reinterpret_cast<const float4*>(z + c);
      const float4 b = *reinterpret_cast<const float4*>(z + c + 4);
      float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = (__expf(g[j] - mx) * inv_se - sm - ((c + j) == y ? (1.0f - eps) : 0.0f)) * inv_nv;
      }
      *reinterpret_cast<uint4*>(dz + c) = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]),
                                                     pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
    }
    for (int c = V8 + tid; c < V; c += 256)
      dz[c] = __float2bfloat16((__expf(z[c] - mx) * inv_se - sm - (c == y ? (1.0f - eps) : 0.0f)) * inv_nv);
  }
}

// ---- fused LM head + cross entropy: combine the per-(row, column half block) partials of the act-5 GEMM
//      (max, sum exp(z - max), sum z) into lse[row] and the label-smoothed loss (modeling_t5.py:1721).
__global__ void __launch_bounds__(256) ce_combine_kernel(const float* __restrict__ stats, int n_slots, const float* __restrict__ zy,
                                                        const long long* __restrict__ labels, const float* __restrict__ n_valid_p,
                                                        float eps, int V, float* __restrict__ lse_out, float* __restrict__ loss_out,
                                                        int M) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* st = stats + (long long)row * n_slots * 3;
  float m = -INFINITY;
  for (int i = lane; i < n_slots; i += 32) m = fmaxf(m, st[i * 3]);
  m = warp_max_f(m);
  float s = 0.f, t = 0.f;
  for (int i = lane; i < n_slots; i += 32) {
    const float mi = st[i * 3];
    if (mi != -INFINITY) s += st[i * 3 + 1] * __expf(mi - m);
    t += st[i * 3 + 2];
  }
  s = warp_sum_f(s);
  t = warp_sum_f(t);
  if (lane == 0) {
    const float lse = m + logf(s);
    lse_out[row] = lse;
    const long long y = labels[row];
    if (y >= 0) {
      const float nll = y < V ? lse - zy[row] : NAN;   // label >= V: caller bug, poison the loss (see cross_entropy_kernel)
      const float smooth = lse - t / (float)V;
      atomicAdd(loss_out, ((1.0f - eps) * nll + eps * smooth) / *n_valid_p);
    }
  }
}

// ---- column sums of a bf16 matrix (bias gradients of the ViT Linears): out[n] += sum_m x[m][n]
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                         float* __restrict__ out, int M, int N, int rows_per_block) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  if (c < N) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += __bfloat162float(x[(long long)r * ld + c]);
    atomicAdd(out + c, s);
  }
}

// ---- fp32 -> bf16 cast with independent row strides (e.g. dQ accumulator -> the q columns of the dQKV matrix)
// HBM-bound: every thread keeps four independent 16-byte loads in flight and the grid is a few CTAs per SM (one load per
// thread in one-shot CTAs ran at ~2 TB/s: the CTA turnover, not the memory system, set the pace).
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd, int M, int N,
                     float scale) {
  pdl_wait();
  pdl_trigger();
  const int n4 = N / 4;
  const long long total = (long long)M * n4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 v[4];
    long long off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      const bool ok = i < total;
      const int r = ok ? (int)(i / n4) : 0, c = ok ? (int)(i % n4) * 4 : 0;
      off[u] = ok ? (long long)r * ldd + c : -1;
      v[u] = ok ? *reinterpret_cast<const float4*>(src + (long long)r * lds + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (off[u] >= 0)
        *reinterpret_cast<uint2*>(dst + off[u]) =
            make_uint2(pack_bf16x2(v[u].x * scale, v[u].y * scale), pack_bf16x2(v[u].z * scale, v[u].w * scale));
  }
}

// ---- copy a bf16 [B,T,C] tensor into rows [row_off, row_off+T) of every batch of a [B,E,C] tensor (memory concat)
__global__ void copy_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int T,
                                      int C, int E, int row_off) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // uint4 (8 elems) index
  const int c8 = C / 8;
  if (i < (long long)B * T * c8) {
    const int c = (int)(i % c8);
    const int t = (int)((i / c8) % T);
    const int b = (int)(i / ((long long)c8 * T));
    reinterpret_cast<uint4*>(dst)[((long long)b * E + row_off + t) * c8 + c] = reinterpret_cast<const uint4*>(src)[i];
  }
}

static inline int cap_grid(long long blocks) {
  const long long cap = (long long)num_sms() * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace vc

using namespace vc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int vc_embed_fwd(const int64_t* ids, const float* table, float* out, int n, int d, int V, uint32_t drop_seed,
                            uint32_t drop_p16, void* stream) {
  VC_CHECK(n > 0 && d % 4 == 0, "vc_embed_fwd: bad dims");
  VC_CUDA(launch_kernel(embed_fwd_kernel, dim3(cap_grid((n + 7) / 8)), dim3(256), 0, ST(stream), (const long long*)ids, table, out, n, d, V, drop_seed,
                                                                  drop_p16, drop_salt_ptr()));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_embed_bwd(const int64_t* ids, const float* dout, float* dtable, int n, int d, int V, uint32_t drop_seed,
                            uint32_t drop_p16, void* stream) {
  VC_CHECK(n > 0 && d % 4 == 0, "vc_embed_bwd: bad dims");
  VC_CUDA(launch_kernel(embed_bwd_kernel, dim3(cap_grid((n + 7) / 8)), dim3(256), 0, ST(stream), (const long long*)ids, dout, dtable, n, d, V, drop_seed,
                                                                  drop_p16, drop_salt_ptr()));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_prepare_targets(const int64_t* out_ids, int64_t* dec_in, int64_t* labels, float* n_valid, int B, int S,
                                  int64_t pad_id, void* stream) {
  VC_CHECK(B > 0 && S > 0, "vc_prepare_targets: bad dims");
  VC_CUDA(cudaMemsetAsync(n_valid, 0, sizeof(float), ST(stream)));
  VC_CUDA(launch_kernel(prepare_targets_kernel, dim3((B * S + 255) / 256), dim3(256), 0, ST(stream), (const long long*)out_ids, (long long*)dec_in,
                                                                     (long long*)labels, n_valid, B, S, pad_id));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_bias_expand(const float* table, const int32_t* lut, float* out, int H, int R, void* stream) {
  VC_CUDA(launch_kernel(bias_expand_kernel, dim3((H * R + 255) / 256), dim3(256), 0, ST(stream), table, lut, out, H, R));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_bias_fold(const float* drel, const int32_t* lut, float* dtable, int H, int R, void* stream) {
  VC_CHECK(H > 0 && R > 0, "vc_bias_fold: bad dims");
  VC_CUDA(launch_kernel(bias_fold_kernel, dim3(H), dim3(256), 0, ST(stream), drel, lut, dtable, H, R));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_add_pos(const float* x, const float* pos, float* out, int B, int T, int C, int P, uint32_t drop_seed,
                          uint32_t drop_p16, void* stream) {
  VC_CHECK(C % 4 == 0, "vc_add_pos: C must be x4");
  const long long total = (long long)B * T * C / 4;
  VC_CUDA(launch_kernel(add_pos_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ST(stream), x, pos, out, B, T, C, P, drop_seed, drop_p16, drop_salt_ptr()));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_add_pos_bwd(const float* dx, float* dpos, int B, int T, int C, int P, uint32_t drop_seed,
                              uint32_t drop_p16, void* stream) {
  VC_CUDA(launch_kernel(add_pos_bwd_kernel, dim3((T * C + 255) / 256), dim3(256), 0, ST(stream), dx, dpos, B, T, C, P, drop_seed, drop_p16, drop_salt_ptr()));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_cross_entropy(const float* logits, int64_t ld, const int64_t* labels, const float* n_valid, float smoothing,
                                float* loss_out, void* dlogits_bf16, int64_t ldd, int n, int V, void* stream) {
  VC_CHECK(n > 0 && ld % 4 == 0 && (dlogits_bf16 == nullptr || ldd % 8 == 0), "vc_cross_entropy: ld alignment");
  VC_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), ST(stream)));
  VC_CUDA(launch_kernel(cross_entropy_kernel, dim3(n), dim3(256), 0, ST(stream), logits, ld, (const long long*)labels, n_valid, smoothing, loss_out,
                                                  (__nv_bfloat16*)dlogits_bf16, ldd, V));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_ce_combine(const float* ce_stats, int n_slots, const float* ce_zy, const int64_t* labels, const float* n_valid,
                             float smoothing, int V, float* lse_out, float* loss_out, int M, void* stream) {
  VC_CHECK(M > 0 && n_slots > 0 && ce_stats && ce_zy && labels && n_valid && lse_out && loss_out, "vc_ce_combine: bad arguments");
  VC_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), ST(stream)));
  VC_CUDA(launch_kernel(ce_combine_kernel, dim3((M + 7) / 8), dim3(256), 0, ST(stream), ce_stats, n_slots, ce_zy,
                        (const long long*)labels, n_valid, smoothing, V, lse_out, loss_out, M));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, void* stream) {
  const int rpb = 64;
  dim3 grid((N + 255) / 256, (M + rpb - 1) / rpb);
  VC_CUDA(launch_kernel(colsum_bf16_kernel, dim3(grid), dim3(256), 0, ST(stream), (const __nv_bfloat16*)x, ld, out, M, N, rpb));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_cast_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int M, int N, float scale,
                                void* stream) {
  VC_CHECK(N % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "vc_cast_f32_bf16: alignment");
  const long long total = (long long)M * (N / 4);
  const long long want = (total + 1023) / 1024, cap = (long long)num_sms() * 8;
  VC_CUDA(launch_kernel(cast_f32_bf16_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), 0, ST(stream), src, lds,
                        (__nv_bfloat16*)dst, ldd, M, N, scale));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_copy_rows_bf16(const void* src, void* dst, int B, int T, int C, int E, int row_off, void* stream) {
  VC_CHECK(C % 8 == 0, "vc_copy_rows_bf16: C must be x8");
  const long long total = (long long)B * T * (C / 8);
  VC_CUDA(launch_kernel(copy_rows_bf16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ST(stream), (const __nv_bfloat16*)src,
                                                                                (__nv_bfloat16*)dst, B, T, C, E, row_off));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
