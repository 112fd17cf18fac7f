// Microbenchmark (design input for the attention kernels): per-SM throughput of tcgen05.ld (TMEM -> registers), of
// MUFU.EX2 and of 16-byte shared-memory stores, as a function of the number of warps.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidchapters_b200/csrc/ptx.cuh"
using namespace vc;

__global__ void __launch_bounds__(512, 1) k_tmem(int iters, int inflight, long long* cyc, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t col0 = ((warp >> 2) * 64) & 511;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (inflight == 1) {
    for (int i = 0; i < iters; ++i) {
      float v[32];
      tmem_ld32(base + ((col0 + (i & 1) * 32) & 511), v);
      tmem_ld_wait();
      acc += v[0] + v[31];
    }
  } else {
    for (int i = 0; i < iters; i += 2) {
      float v[32], w[32];
      tmem_ld32(base + (col0 & 511), v);
      tmem_ld32(base + ((col0 + 32) & 511), w);
      tmem_ld_wait();
      acc += v[0] + v[31] + w[0] + w[31];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

__global__ void __launch_bounds__(512, 1) k_ex2(int iters, long long* cyc, float* sink) {
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fast_exp2(x[j]) - 1.0f;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(512, 1) k_sts(int iters, long long* cyc, float* sink) {
  extern __shared__ uint4 sm[];
  const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[(threadIdx.x + j * blockDim.x) & 4095] = v;   // conflict-free 16-byte stores
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (sm[threadIdx.x].x == 0xdeadbeef) sink[0] = 1.f;
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  long long h[148];
  const int iters = 4096;
  for (int inflight = 1; inflight <= 2; ++inflight)
    for (int warps = 4; warps <= 16; warps *= 2) {
      k_tmem<<<148, warps * 32>>>(iters, inflight, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("tcgen05.ld 32x32b.x32  warps=%2d inflight=%d : %.0f cycles for %d loads/warp -> %.1f B/clk/SM (%.1f clk per 4 KB load per warp)\n",
             warps, inflight, c, iters, (double)warps * iters * 4096.0 / c, c / iters);
    }
  for (int warps = 4; warps <= 16; warps *= 2) {
    k_ex2<<<148, warps * 32>>>(iters, cyc, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("MUFU.EX2 (+FADD)       warps=%2d : %.2f ex2/clk/SM\n", warps, (double)warps * 32 * iters * 8 / c);
  }
  cudaFuncSetAttribute(k_sts, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int warps = 4; warps <= 16; warps *= 2) {
    k_sts<<<148, warps * 32, 65536>>>(iters, cyc, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("STS.128                warps=%2d : %.1f B/clk/SM\n", warps, (double)warps * 32 * iters * 8 * 16 / c);
  }
  return 0;
}
