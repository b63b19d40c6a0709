// Library plumbing: error reporting, device query, launch accounting.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "et_common.cuh"
#include "et_tma.cuh"

namespace et {

static thread_local char g_err[512] = "";
static std::atomic<int> g_tune[ET_TUNE_COUNT];

int tune_get(int knob) { return (knob >= 0 && knob < ET_TUNE_COUNT) ? g_tune[knob].load(std::memory_order_relaxed) : 0; }
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return ET_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// cuTensorMapEncodeTiled is resolved through the runtime so that the library does not link libcuda.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encoder() {
  static std::atomic<encode_tiled_fn> cached{nullptr};
  encode_tiled_fn fn = cached.load(std::memory_order_acquire);
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<encode_tiled_fn>(p);
  cached.store(fn, std::memory_order_release);
  return fn;
}

int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint32_t cols, uint64_t row_pitch_bytes,
                       uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  encode_tiled_fn enc = get_encoder();
  if (!enc) return fail(ET_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ET_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ET_OK;
}

}  // namespace et

extern "C" {

int et_version(void) { return ET_B200_VERSION; }

int et_tune(int knob, int value) {
  if (knob < 0 || knob >= ET_TUNE_COUNT) return et::fail(ET_ERR_BADARG, "et_tune: unknown knob %d", knob);
  et::g_tune[knob].store(value, std::memory_order_relaxed);
  return ET_OK;
}

const char* et_last_error(void) { return et::g_err; }

int64_t et_launch_count(void) { return et::g_launches.load(std::memory_order_relaxed); }

int et_memcpy_2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                       et_stream_t stream) {
  if (rows == 0 || width_bytes == 0) return ET_OK;
  if (!dst || !src || dst_pitch < width_bytes || src_pitch < width_bytes)
    return et::fail(ET_ERR_BADARG, "et_memcpy_2d_async: bad pointers or pitches");
  cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDefault, et::as_stream(stream));
  if (e != cudaSuccess) return et::fail(ET_ERR_CUDA, "cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
  return ET_OK;
}

int et_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return et::fail(ET_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int sms = 0, maj = 0, min = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return et::fail(ET_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return ET_OK;
}

}  // extern "C"
