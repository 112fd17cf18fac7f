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
extern "C" int vc_step_advance(int32_t* pos_dev, void* stream) {
  VC_CUDA(launch_kernel(step_advance_kernel, dim3(1), dim3(1), 0, ST(stream), pos_dev));
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}
