// metrics.cu — image-quality metrics on the device: PSNR and SSIM of a dehazed batch against
// the clear batch without the per-image device->host copy + skimage call of the reference's evaluation loop
// (evaluation/metrics.py:13-36, training/train_dehazing.py:146-159, evaluation/evaluate.py:158-168).
//
//   PSNR_i = 10 log10(1 / mean_{c,y,x} (p - t)^2)                      (skimage peak_signal_noise_ratio, data_range = 1)
//   SSIM_i = skimage structural_similarity(gray_t, gray_p, data_range = 1) with its defaults: gray = mean over the
//            colour channels, 7x7 uniform window, sample covariance (x 49/48), K1 = 0.01, K2 = 0.03, mean of the SSIM
//            map over the interior [3, H-3) x [3, W-3).
#include "adb_ptx.cuh"
#include "adb_host.h"
#include <algorithm>

namespace {

constexpr int kTW = 32, kTH = 16, kR = 3;

// acc[i][0] += sum (p-t)^2 over 3hw ; acc[i][1] += sum of the SSIM map over the interior
__global__ void image_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int h, int w,
                                     double* __restrict__ acc) {
  __shared__ float s_p[kTH + 2 * kR][kTW + 2 * kR];
  __shared__ float s_t[kTH + 2 * kR][kTW + 2 * kR];
  __shared__ float s_red[2][(kTW * kTH) / 32];
  const int img = blockIdx.z;
  const int tid = threadIdx.y * kTW + threadIdx.x;
  const int x0 = blockIdx.x * kTW - kR, y0 = blockIdx.y * kTH - kR;
  const size_t plane = (size_t)h * w;
  const float* p = pred + (size_t)img * 3 * plane;
  const float* t = tgt + (size_t)img * 3 * plane;
  for (int i = tid; i < (kTH + 2 * kR) * (kTW + 2 * kR); i += kTW * kTH) {
    const int ty = i / (kTW + 2 * kR), tx = i - ty * (kTW + 2 * kR);
    const int yy = y0 + ty, xx = x0 + tx;
    float gp = 0.f, gt = 0.f;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const size_t o = (size_t)yy * w + xx;
      gp = (__ldg(p + o) + __ldg(p + plane + o) + __ldg(p + 2 * plane + o)) * (1.f / 3.f);
      gt = (__ldg(t + o) + __ldg(t + plane + o) + __ldg(t + 2 * plane + o)) * (1.f / 3.f);
    }
    s_p[ty][tx] = gp;
    s_t[ty][tx] = gt;
  }
  __syncthreads();
  const int px = blockIdx.x * kTW + threadIdx.x, py = blockIdx.y * kTH + threadIdx.y;
  float se = 0.f, ss = 0.f;
  if (px < w && py < h) {
    const size_t o = (size_t)py * w + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float d = __ldg(p + c * plane + o) - __ldg(t + c * plane + o); se = fmaf(d, d, se); }
    if (px >= kR && px < w - kR && py >= kR && py < h - kR) {
      float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
      for (int dy = 0; dy < 7; ++dy)
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const float a = s_t[threadIdx.y + dy][threadIdx.x + dx], b = s_p[threadIdx.y + dy][threadIdx.x + dx];
          sx += a; sy += b; sxx = fmaf(a, a, sxx); syy = fmaf(b, b, syy); sxy = fmaf(a, b, sxy);
        }
      const float inv = 1.f / 49.f, cn = 49.f / 48.f;
      const float ux = sx * inv, uy = sy * inv;
      const float vx = cn * (sxx * inv - ux * ux), vy = cn * (syy * inv - uy * uy), vxy = cn * (sxy * inv - ux * uy);
      const float C1 = 1e-4f, C2 = 9e-4f;
      ss = ((2.f * ux * uy + C1) * (2.f * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if (lane == 0) { s_red[0][warp] = se; s_red[1][warp] = ss; }
  __syncthreads();
  if (tid < 2) {
    double v = 0.0;
    for (int k = 0; k < (kTW * kTH) / 32; ++k) v += (double)s_red[tid][k];
    atomicAdd(acc + (size_t)img * 2 + tid, v);
  }
}

__global__ void image_metrics_finish_kernel(const double* __restrict__ acc, int n, double inv_mse_n, double inv_ssim_n,
                                            float* __restrict__ psnr, float* __restrict__ ssim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mse = acc[(size_t)i * 2] * inv_mse_n;
  psnr[i] = mse > 0.0 ? (float)(10.0 * log10(1.0 / mse)) : INFINITY;
  ssim[i] = (float)(acc[(size_t)i * 2 + 1] * inv_ssim_n);
}

}  // namespace

extern "C" int adb_image_metrics(const float* pred, const float* target, int32_t n, int32_t h, int32_t w, double* scratch /*[2n]*/,
                                 float* psnr /*[n]*/, float* ssim /*[n]*/, void* stream) {
  ADB_REQUIRE(pred && target && scratch && psnr && ssim && n > 0, "adb_image_metrics: null pointer");
  ADB_REQUIRE(h >= 7 && w >= 7, "adb_image_metrics: SSIM needs images of at least 7x7 (got %dx%d)", h, w);
  cudaStream_t st = (cudaStream_t)stream;
  ADB_CUDA_OK(cudaMemsetAsync(scratch, 0, (size_t)n * 2 * sizeof(double), st));
  dim3 block(kTW, kTH), grid((w + kTW - 1) / kTW, (h + kTH - 1) / kTH, n);
  image_metrics_kernel<<<grid, block, 0, st>>>(pred, target, h, w, scratch);
  ADB_CUDA_OK(cudaGetLastError());
  image_metrics_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(scratch, n, 1.0 / (3.0 * h * w), 1.0 / ((double)(h - 6) * (w - 6)), psnr, ssim);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}
