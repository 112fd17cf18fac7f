// Incremental-decoding helpers (HBM-bound): KV-cache append, greedy next-token selection, step counter.
//
// Replaces, for `Vid2Seq.generate` with num_beams=1 (reference model/vid2seq.py:150-162 -> HF-4.28 greedy_search, which
// is third-party code; semantics restated in SURVEY.md §8c): per step the new self-attention K/V row is appended to the
// cache (modeling_t5.py:511-515 torch.cat), next = argmax(logits); sequences that already produced eos (1) emit pad (0).
// The position lives in DEVICE memory so that one decode step is a fixed CUDA graph replayed max_new_tokens times.
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace vc {

// cache[b, *pos, :] = src[b, :]   (src [B][C] bf16 with row stride lds; cache [B][cap][C])
__global__ void kv_append_kernel(const __nv_bfloat16* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ cache, int B,
                                 int cap, int C, const int* __restrict__ pos_dev) {
  pdl_wait();
  pdl_trigger();
  const int pos = *pos_dev;
  const int c8 = C / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * c8 && pos < cap) {
    const int b = i / c8, c = i % c8;
    reinterpret_cast<uint4*>(cache + ((long long)b * cap + pos) * C)[c] =
        reinterpret_cast<const uint4*>(src + (long long)b * lds)[c];
  }
}

// One block per sequence: next = done ? pad : argmax_v logits[b, v] (first maximum, like torch.argmax);
// seq[b, pos + 1] = next; ids_out[b] = next (input of the next step); done |= (next == eos).
__global__ void __launch_bounds__(256)
greedy_next_kernel(const float* __restrict__ logits, long long ld, int V, unsigned char* __restrict__ done,
                   long long* __restrict__ ids_out, long long* __restrict__ seq, int seq_ld, const int* __restrict__ pos_dev,
                   long long eos_id, long long pad_id) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  const float* z = logits + (long long)b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float x = z[v];
    if (x > best) { best = x; bi = v; }   // strided scan keeps the smallest index per thread among equals
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
    const int pos = *pos_dev;
    long long next = done[b] ? pad_id : (long long)bi;
    if (next == eos_id) done[b] = 1;
    ids_out[b] = next;
    if (pos + 1 < seq_ld) seq[(long long)b * seq_ld + pos + 1] = next;
  }
}

// ---- beam search (vid2seq.py:150-162 with num_beams > 1 -> HF-4.28 beam_search): per batch item, the top K2 = 2*nb
// candidates of  log_softmax(logits[b*nb + r, :])[v] + beam_scores[b*nb + r]  over (r, v), sorted descending (ties: lower
// flat index r*V + v first).  One block per batch item: row log-sum-exps, a sorted top-K2 list per thread, then K2 rounds
// of block arg-max over the 256 x K2 shortlisted candidates.
constexpr int kBeamK2Max = 16;
__global__ void __launch_bounds__(256)
beam_topk_kernel(const float* __restrict__ logits, long long ld, int V, const float* __restrict__ beam_scores, int nb, int k2,
                 float* __restrict__ out_scores, int* __restrict__ out_tokens, int* __restrict__ out_beams) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_red[8];
  __shared__ float s_lse[kBeamK2Max / 2];
  __shared__ float s_cs[256 * kBeamK2Max];
  __shared__ int s_ci[256 * kBeamK2Max];
  __shared__ float s_bv[8];
  __shared__ int s_bi[8], s_bslot[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // ---- row log-sum-exp
  for (int r = 0; r < nb; ++r) {
    const float* z = logits + ((long long)b * nb + r) * ld;
    float m = -INFINITY;
    for (int v = tid; v < V; v += 256) m = fmaxf(m, z[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[wid] = m;
    __syncthreads();
    m = s_red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    __syncthreads();
    float e = 0.f;
    for (int v = tid; v < V; v += 256) e += expf(z[v] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) s_red[wid] = e;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_red[w];
      s_lse[r] = m + logf(t);
    }
    __syncthreads();
  }
  // ---- per-thread sorted shortlist (descending; equal scores keep the lower flat index first)
  float ls[kBeamK2Max];
  int li[kBeamK2Max];
#pragma unroll
  for (int j = 0; j < kBeamK2Max; ++j) { ls[j] = -INFINITY; li[j] = 0x7fffffff; }
  for (int r = 0; r < nb; ++r) {
    const float* z = logits + ((long long)b * nb + r) * ld;
    const float add = beam_scores[b * nb + r] - s_lse[r];
    for (int v = tid; v < V; v += 256) {
      float c = z[v] + add;
      int ci = r * V + v;
      if (c > ls[kBeamK2Max - 1] || (c == ls[kBeamK2Max - 1] && ci < li[kBeamK2Max - 1])) {
#pragma unroll
        for (int j = 0; j < kBeamK2Max; ++j) {
          const bool better = c > ls[j] || (c == ls[j] && ci < li[j]);
          const float ts = ls[j]; const int ti = li[j];
          ls[j] = better ? c : ts; li[j] = better ? ci : ti;
          c = better ? ts : c; ci = better ? ti : ci;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kBeamK2Max; ++j) { s_cs[tid * kBeamK2Max + j] = ls[j]; s_ci[tid * kBeamK2Max + j] = li[j]; }
  __syncthreads();
  // ---- K2 rounds of block arg-max over the shortlist
  for (int k = 0; k < k2; ++k) {
    float bv = -INFINITY; int bi = 0x7fffffff, bslot = -1;
    for (int c = tid; c < 256 * kBeamK2Max; c += 256) {
      const float x = s_cs[c]; const int xi = s_ci[c];
      if (xi != 0x7fffffff && (x > bv || (x == bv && xi < bi))) { bv = x; bi = xi; bslot = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int os = __shfl_xor_sync(0xffffffffu, bslot, o);
      if (os >= 0 && (bslot < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bslot = os; }
    }
    if (lane == 0) { s_bv[wid] = bv; s_bi[wid] = bi; s_bslot[wid] = bslot; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_bslot[w] >= 0 && (bslot < 0 || s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi))) { bv = s_bv[w]; bi = s_bi[w]; bslot = s_bslot[w]; }
      out_scores[b * k2 + k] = bv;
      out_tokens[b * k2 + k] = bslot >= 0 ? bi % V : 0;
      out_beams[b * k2 + k] = bslot >= 0 ? bi / V : 0;
      if (bslot >= 0) s_ci[bslot] = 0x7fffffff;   // taken
    }
    __syncthreads();
  }
}

// cache rows [0, n) of every sequence follow their beam: dst[b] = src[beam_idx[b]] (HF _reorder_cache, modeling_t5.py:1771-1793)
__global__ void kv_reorder_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                  const int* __restrict__ beam_idx, int Bn, int cap, int C, int n) {
  pdl_wait();
  pdl_trigger();
  const int c8 = C / 8;
  const long long total = (long long)Bn * n * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    const long long t = i / c8;
    const int row = (int)(t % n), b = (int)(t / n);
    reinterpret_cast<uint4*>(dst + ((long long)b * cap + row) * C)[c] =
        reinterpret_cast<const uint4*>(src + ((long long)beam_idx[b] * cap + row) * C)[c];
  }
}

__global__ void step_advance_kernel(int* pos_dev) {
  pdl_wait();
  pdl_trigger(); *pos_dev += 1; }

}  // namespace vc

using namespace vc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int vc_kv_append(const void* src, int64_t lds, void* cache, int B, int cap, int C, const int32_t* pos_dev,
                            void* stream) {
  VC_CHECK(B > 0 && C % 8 == 0 && lds % 8 == 0, "vc_kv_append: C and lds must be multiples of 8");
  VC_CUDA(launch_kernel(kv_append_kernel, dim3((B * (C / 8) + 255) / 256), dim3(256), 0, ST(stream), (const __nv_bfloat16*)src, lds, (__nv_bfloat16*)cache, B,
                                                                     cap, C, pos_dev));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_greedy_next(const float* logits, int64_t ld, int V, uint8_t* done, int64_t* ids_out, int64_t* seq, int seq_ld,
                              const int32_t* pos_dev, int64_t eos_id, int64_t pad_id, int B, void* stream) {
  VC_CHECK(B > 0 && V > 0, "vc_greedy_next: bad dims");
  VC_CUDA(launch_kernel(greedy_next_kernel, dim3(B), dim3(256), 0, ST(stream), logits, ld, V, done, (long long*)ids_out, (long long*)seq, seq_ld, pos_dev,
                                                eos_id, pad_id));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_beam_topk(const float* logits, int64_t ld, int V, const float* beam_scores, int num_beams, int B,
                            float* out_scores, int32_t* out_tokens, int32_t* out_beams, void* stream) {
  VC_CHECK(B > 0 && V > 0 && num_beams >= 1 && 2 * num_beams <= kBeamK2Max, "vc_beam_topk: 1 <= num_beams <= %d", kBeamK2Max / 2);
  VC_CHECK((long long)num_beams * V < 0x7fffffffLL, "vc_beam_topk: num_beams * V overflows");
  VC_CUDA(launch_kernel(beam_topk_kernel, dim3(B), dim3(256), 0, ST(stream), logits, (long long)ld, V, beam_scores, num_beams,
                        2 * num_beams, out_scores, (int*)out_tokens, (int*)out_beams));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_kv_reorder(const void* src, void* dst, const int32_t* beam_idx, int Bn, int cap, int C, int n, void* stream) {
  VC_CHECK(Bn > 0 && C % 8 == 0 && n >= 0 && n <= cap && src != dst, "vc_kv_reorder: bad args");
  if (n == 0) return VC_OK;
  const long long total = (long long)Bn * n * (C / 8);
  const int grid = (int)((total + 255) / 256 < (long long)num_sms() * 16 ? (total + 255) / 256 : (long long)num_sms() * 16);
  VC_CUDA(launch_kernel(kv_reorder_kernel, dim3(grid), dim3(256), 0, ST(stream), (const __nv_bfloat16*)src, (__nv_bfloat16*)dst,
                        (const int*)beam_idx, Bn, cap, C, n));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
extern "C" int vc_step_advance(int32_t* pos_dev, void* stream) {
  VC_CUDA(launch_kernel(step_advance_kernel, dim3(1), dim3(1), 0, ST(stream), pos_dev));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
