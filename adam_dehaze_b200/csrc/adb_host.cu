// adb_host.cu — error text, device query, TMA descriptor encoder lookup, kernel error flag.
#include "adb_host.h"
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace adbh {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

static const int kMaxDev = 16;
static DeviceInfo g_dev[kMaxDev];
static int* g_flag[kMaxDev];
static std::mutex g_mu;

int device_info(DeviceInfo* out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= kMaxDev)
    return fail(ADB_ERR_NO_DEVICE, "no CUDA device available: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_dev[dev].ok) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return fail(ADB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    g_dev[dev].sm_count = p.multiProcessorCount;
    g_dev[dev].cc_major = p.major;
    g_dev[dev].cc_minor = p.minor;
    g_dev[dev].max_smem_optin = (int)p.sharedMemPerBlockOptin;
    g_dev[dev].ok = 1;
  }
  *out = g_dev[dev];
  return ADB_OK;
}

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_span_bytes) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail(ADB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_span_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_span_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                        : swizzle_span_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(ADB_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] swz %d",
                (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
                rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0, swizzle_span_bytes);
  }
  return ADB_OK;
}

int* kernel_err_flag() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_flag[dev]) {
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(int));
    g_flag[dev] = p;
  }
  return g_flag[dev];
}

}  // namespace adbh

extern "C" {

const char* adb_last_error(void) { return adbh::g_err; }

int adb_version(void) { return 100; }

int adb_device_check(void) {
  adbh::DeviceInfo di;
  int st = adbh::device_info(&di);
  if (st != ADB_OK) return st;
  if (di.cc_major != 10) return adbh::fail(ADB_ERR_NO_DEVICE, "device is sm_%d%d, this library is sm_100a only", di.cc_major, di.cc_minor);
  if (!adbh::encode_tiled()) return adbh::fail(ADB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  return ADB_OK;
}

int adb_kernel_error_flag(void) {
  int* f = adbh::kernel_err_flag();
  if (!f) return adbh::fail(ADB_ERR_NO_DEVICE, "no device");
  int v = 0;
  cudaError_t e = cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return adbh::fail(ADB_ERR_CUDA, "error flag read: %s", cudaGetErrorString(e));
  if (v != 0) {
    cudaMemset(f, 0, sizeof(int));
    return adbh::fail(ADB_ERR_KERNEL, "device-side error flag = %d (1-30: a pipeline wait of a conv kernel timed out; 31: class label outside [0, classes) in adb_ce_fwd_bwd)", v);
  }
  return ADB_OK;
}

}  // extern "C"
