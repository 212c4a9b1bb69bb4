// route.cu — device-side routing: argmax over the HDEN logits and a stable 3-way bucketing, one launch, no host sync.
// Replaces models/routing.py:40-61 (torch.argmax, the three `intensity == k` masks, and the x[mask] gathers whose
// nonzero() calls synchronise the host): bucket k lists, in ascending order, the batch rows routed to branch k.
#include "adb_ptx.cuh"
#include "adb_host.h"

namespace {

constexpr int kRouteThreads = 1024;

// torch.argmax semantics: the first maximal element wins; NaN compares greater than everything (first NaN wins).
__device__ __forceinline__ int argmax_row(const float* l, int classes) {
  int best = 0;
  float bv = l[0];
  for (int k = 1; k < classes; ++k) {
    const float v = l[k];
    const bool better = (bv != bv) ? false : ((v != v) ? true : (v > bv));
    if (better) { best = k; bv = v; }
  }
  return best;
}

__global__ void __launch_bounds__(kRouteThreads, 1)
route_kernel(const float* __restrict__ logits, const long long* __restrict__ intensity_in, int b, int classes,
             long long* __restrict__ intensity, uint8_t* __restrict__ masks, int* __restrict__ bucket_index,
             int* __restrict__ bucket_count) {
  __shared__ int warp_cnt[3][kRouteThreads / 32];
  __shared__ int base[3];
  __shared__ int chunk_base[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 3) base[threadIdx.x] = 0;
  __syncthreads();
  for (int start = 0; start < b; start += kRouteThreads) {
    const int i = start + threadIdx.x;
    int cls = -1;
    if (i < b) {
      cls = intensity_in ? (int)intensity_in[i] : argmax_row(logits + (size_t)i * classes, classes);
      intensity[i] = cls;
#pragma unroll
      for (int k = 0; k < 3; ++k) masks[(size_t)k * b + i] = (cls == k) ? 1 : 0;
    }
    int pos_in_warp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned m = __ballot_sync(0xffffffffu, cls == k);
      pos_in_warp[k] = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) warp_cnt[k][warp] = __popc(m);
    }
    __syncthreads();
    // exclusive scan of the 32 warp counts, one warp per class
    if (warp < 3) {
      const int v = warp_cnt[warp][lane];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      warp_cnt[warp][lane] = inc - v;
      const int tot = __shfl_sync(0xffffffffu, inc, 31);
      __syncwarp();
      if (lane == 0) { const int old = base[warp]; chunk_base[warp] = old; base[warp] = old + tot; }
    }
    __syncthreads();
    if (cls >= 0 && cls < 3) {
      bucket_index[(size_t)cls * b + chunk_base[cls] + warp_cnt[cls][warp] + pos_in_warp[cls]] = i;
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) bucket_count[threadIdx.x] = base[threadIdx.x];
}

// rows whose class id is outside {0,1,2} are written by no branch: the reference leaves them at zeros_like(x)
// (routing.py:31,55-61).  Every other row is overwritten in full by its branch's image epilogue, so only these are cleared.
__global__ void zero_unrouted_kernel(float* __restrict__ out, const long long* __restrict__ intensity, long long row_elems) {
  const long long cls = intensity[blockIdx.y];
  if (cls >= 0 && cls < 3) return;
  float4* row = reinterpret_cast<float4*>(out + (size_t)blockIdx.y * row_elems);
  const long long n4 = row_elems >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x == 0 && threadIdx.x < (row_elems & 3)) out[(size_t)blockIdx.y * row_elems + (n4 << 2) + threadIdx.x] = 0.f;
}

}  // namespace

extern "C" int adb_zero_unrouted(float* out, const int64_t* intensity, int32_t b, int64_t row_elems, void* stream) {
  ADB_REQUIRE(out && intensity && b > 0 && b <= 65535 && row_elems > 0, "adb_zero_unrouted: bad arguments");
  ADB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (row_elems % 4 == 0 || b == 1), "adb_zero_unrouted: rows must be 16-byte aligned");
  zero_unrouted_kernel<<<dim3(64, (unsigned)b), 256, 0, (cudaStream_t)stream>>>(out, reinterpret_cast<const long long*>(intensity), row_elems);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

extern "C" int adb_route(const float* logits, const int64_t* intensity_in, int32_t b, int32_t classes,
                         int64_t* intensity, uint8_t* masks, int32_t* bucket_index, int32_t* bucket_count, void* stream) {
  ADB_REQUIRE((logits || intensity_in) && intensity && masks && bucket_index && bucket_count, "adb_route: null pointer");
  ADB_REQUIRE(b > 0 && classes >= 1 && classes <= 64, "adb_route: bad batch %d / classes %d", b, classes);
  route_kernel<<<1, kRouteThreads, 0, (cudaStream_t)stream>>>(logits, reinterpret_cast<const long long*>(intensity_in), b,
                                                             classes, reinterpret_cast<long long*>(intensity), masks,
                                                             bucket_index, bucket_count);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}
