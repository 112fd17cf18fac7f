// Error reporting, device check and TMA tensor-map encoding shared by all entry points.
#include <stdarg.h>
#include <string.h>

#include "common.h"

namespace vc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VC_OK;
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return VC_ERR_CUDA;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VIDCHAP_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VC_ERR_CUDA;
  }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu stride0 %llu box %u,%u base %p", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)strides_bytes[0], box[0], box[1], base);
    return VC_ERR_CUDA;
  }
  return VC_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                 uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  return encode(out, base, 2, dims, strides, box);
}

int make_tmap_2d_ex(CUtensorMap* out, const void* base, int elt_bytes, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * (uint64_t)elt_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, elt_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(ex) failed (%d): elt %d dims %llu,%llu stride %llu box %u,%u swz %d base %p", (int)r,
              elt_bytes, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)strides[0], box_inner,
              box_outer, swizzle_bytes, base);
    return VC_ERR_CUDA;
  }
  return VC_OK;
}

int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2, uint64_t s1_elems,
                 uint64_t s2_elems, uint32_t box_inner, uint32_t box_d1) {
  cuuint64_t dims[3] = {inner, d1, d2};
  cuuint64_t strides[2] = {s1_elems * 2, s2_elems * 2};
  cuuint32_t box[3] = {box_inner, box_d1, 1};
  return encode(out, base, 3, dims, strides, box);
}

static const uint32_t* g_drop_salt = nullptr;
const uint32_t* drop_salt_ptr() { return g_drop_salt; }

static long long* g_trace = nullptr;
long long* debug_trace_ptr() { return g_trace; }

}  // namespace vc

extern "C" int vc_debug_set_trace(void* dev_ptr) {
  vc::g_trace = reinterpret_cast<long long*>(dev_ptr);
  return VC_OK;
}
extern "C" int vc_set_dropout_salt(const uint32_t* dev_ptr) {
  vc::g_drop_salt = dev_ptr;
  return VC_OK;
}
extern "C" int vc_version(void) { return 9; }
extern "C" const char* vc_last_error(void) { return vc::g_err; }
extern "C" int vc_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (vc::check_cuda(cudaGetDevice(&dev), "cudaGetDevice") != VC_OK) return VC_ERR_CUDA;
  if (vc::check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties") != VC_OK) return VC_ERR_CUDA;
  if (prop.major != 10) {
    vc::set_error("libvidchap needs an sm_100 device (B200); found sm_%d%d (%s)", prop.major, prop.minor, prop.name);
    return VC_ERR_INVALID;
  }
  return VC_OK;
}
