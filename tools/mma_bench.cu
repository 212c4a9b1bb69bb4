// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, SS mode) as a function of M, N and the operand
// swizzle span.  Operands are whatever is in shared memory; only the issue/execute rate is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu && ./mma_bench
#include "../adam_dehaze_b200/csrc/adb_ptx.cuh"
#include <cstdio>
using namespace adb;

__device__ __forceinline__ uint32_t idesc_mn(uint32_t M, uint32_t N) { return make_idesc_bf16(M, N); }

__global__ void __launch_bounds__(128, 1) bench(int M, int N, int row_bytes, int reps, int distinct_b, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 0) {
    const uint64_t hi = make_kmajor_desc(0, row_bytes);
    const uint32_t a_addr = base, b_addr = base + 64 * 1024;
    const uint32_t idesc = idesc_mn(M, N);
    const int ksteps = row_bytes / 32;
    long long t0 = clock64();
    if (lane == 0) {
      for (int r = 0; r < reps; ++r) {
        const uint64_t a0 = hi | (uint64_t)(((a_addr + (r & 3) * 16384) & 0x3FFFFu) >> 4);
        const uint64_t b0 = hi | (uint64_t)(((b_addr + (distinct_b ? (r & 3) * 32768 : 0)) & 0x3FFFFu) >> 4);
        for (int kk = 0; kk < ksteps; ++kk) umma_bf16(tmem, a0 + kk * 2, b0 + kk * 2, idesc, 1u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ms[] = {128, 64};
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  const int RBs[] = {128, 64, 32};
  for (int grid : {1, 148})
    for (int M : Ms) for (int rb : RBs) for (int N : Ns) {
      const int reps = 256;
      bench<<<grid, 128, 200 * 1024>>>(M, N, rb, reps, 1, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const int nmma = reps * (rb / 32);
      printf("grid %3d M %3d N %3d swz %3d : issue %6.1f cyc/mma, complete %6.1f cyc/mma  (floor %5.1f) %s\n", grid, M, N, rb,
             (double)h[0] / nmma, (double)h[1] / nmma, (M < 128 ? 128 : M) * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
