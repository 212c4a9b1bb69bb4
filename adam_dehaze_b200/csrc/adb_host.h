// adb_host.h — host-side helpers shared by the C-ABI translation units (error text, device info, TMA encoder).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/adb200.h"

namespace adbh {

void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define ADB_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return adbh::fail(ADB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define ADB_REQUIRE(cond, ...)                                        \
  do {                                                                \
    if (!(cond)) return adbh::fail(ADB_ERR_INVALID, __VA_ARGS__);     \
  } while (0)

struct DeviceInfo {
  int ok;        // 1 when queried
  int sm_count;
  int cc_major, cc_minor;
  int max_smem_optin;
};
int device_info(DeviceInfo* out);  // for the current device; cached per device id

// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (libcuda is never linked directly, so the
// library loads on a machine without a driver and only fails when a kernel is requested).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled();

// bf16 tensor map; dims/strides innermost-first; strides[i] is the byte stride of dim i+1.
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_span_bytes);

int* kernel_err_flag();  // device int, zero-initialised, per device

}  // namespace adbh
