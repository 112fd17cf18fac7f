// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/vidchap.h"

namespace vc {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define VC_CHECK(cond, ...)            \
  do {                                 \
    if (!(cond)) {                     \
      vc::set_error(__VA_ARGS__);      \
      return VC_ERR_INVALID;           \
    }                                  \
  } while (0)
#define VC_CUDA(call)                                      \
  do {                                                     \
    int _s = vc::check_cuda((call), #call);                \
    if (_s != VC_OK) return _s;                            \
  } while (0)

// bf16 tensor maps, SWIZZLE_128B, zero OOB fill.  Box inner extent is always 64 elements (128 B).
//  2D: tensor [outer][inner] with row stride `row_stride_elems`.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                 uint32_t box_inner, uint32_t box_outer);
//  3D: tensor [d2][d1][inner] with element strides s1 (between d1 rows) and s2 (between d2 slabs).
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2, uint64_t s1_elems,
                 uint64_t s2_elems, uint32_t box_inner, uint32_t box_d1);

// General 2-D map: elt_bytes 2 (bf16) or 4 (fp32); swizzle_bytes 0/32/64/128.
int make_tmap_2d_ex(CUtensorMap* out, const void* base, int elt_bytes, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);

int num_sms();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: a process that launches on cuda:1 after cuda:0 must
// set it again there (a single static flag made the first >48 KB launch on the second device fail with "invalid argument").
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Programmatic dependent launch for every kernel of the library (VIDCHAP_PDL=0 turns it off for A/B runs); see
// ptx.cuh::pdl_wait.  Inside stream capture the attribute becomes a programmatic dependency edge of the CUDA graph.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Device pointer to a 32-bit salt XORed into every dropout seed (vc_set_dropout_salt); lets a captured CUDA graph draw
// fresh masks on every replay.  nullptr = no salt.
const uint32_t* drop_salt_ptr();
// Debug only (vc_debug_set_trace): device buffer of >= 2048 int64 that the attention kernels fill with a clock64
// timeline of one CTA; nullptr (default) = off.
long long* debug_trace_ptr();

}  // namespace vc
