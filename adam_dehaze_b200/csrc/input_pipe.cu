// input_pipe.cu — the per-sample input transform of the reference loader on the device.
//
// data/dataset.py:73-99 does, on CPU workers and three times per sample: cv2.imread (BGR uint8 HWC) -> cv2.cvtColor BGR2RGB ->
// cv2.resize(img_size) when the size differs -> transforms.ToTensor() (uint8 HWC -> float32 CHW / 255); the training
// split adds RandomHorizontal/VerticalFlip on the tensor (dataset.py:58-63).  Here the decoded uint8 HWC image is uploaded
// as it is (a quarter of the bytes of the float tensor) and ONE kernel produces the NCHW fp32 batch the models consume:
// channel swap, OpenCV's INTER_LINEAR for uint8 reproduced bit for bit, /255, optional flips.
//
// cv2.resize, INTER_LINEAR, 8-bit (opencv/modules/imgproc/src/resize.cpp, 4.x), as restated and pinned in
// oracle/input_oracle.py:
//   * an exact 2x downscale in both axes takes the INTER_AREA fast path: (a + b + c + d + 2) >> 2;
//   * otherwise, per axis: f = (float)((d + 0.5) * scale - 0.5) (double product, rounded to float), s = floor(f), frac = f - s;
//     coefficients are shorts: rint((1 - frac) * 2048), rint(frac * 2048);
//     columns: s < 0 -> s = 0, frac = 0;  s >= n-1 -> s = n-1, frac = 0;      rows: only the indices are clamped;
//     horizontal pass in int: r = S[x0]*a0 + S[x1]*a1;   vertical: ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2.
#include "adb_host.h"

namespace {

struct AxisTap { int i0, i1, c0, c1; };

__device__ __forceinline__ AxisTap axis_tap(int d, int dn, int sn, bool vertical) {
  // (explicit round-to-nearest intrinsics: no fused multiply-add contraction, the host code has none)
  const double scale = __ddiv_rn((double)sn, (double)dn);
  const float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = (int)floorf(f);
  float frac = __fsub_rn(f, (float)s);
  if (!vertical) {
    if (s < 0) { s = 0; frac = 0.f; }
    if (s >= sn - 1) { s = sn - 1; frac = 0.f; }
  }
  AxisTap t;
  t.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, frac), 2048.f));
  t.c1 = __float2int_rn(__fmul_rn(frac, 2048.f));
  t.i0 = min(max(s, 0), sn - 1);
  t.i1 = min(max(s + 1, 0), sn - 1);
  return t;
}

__global__ void image_u8_to_f32_kernel(const uint8_t* __restrict__ src, int n, int hs, int ws, long long src_stride, int bgr,
                                       const uint8_t* __restrict__ flips, float* __restrict__ dst, int hd, int wd) {
  const long long total = (long long)n * hd * wd;
  const bool same = hs == hd && ws == wd;
  const bool area2 = hs == 2 * hd && ws == 2 * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wd);
    const int y = (int)((i / wd) % hd);
    const int img = (int)(i / ((long long)wd * hd));
    const uint8_t* s = src + (size_t)img * src_stride;
    int v[3];
    if (same) {
      const uint8_t* p = s + ((size_t)y * ws + x) * 3;
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    } else if (area2) {
      const uint8_t* p0 = s + ((size_t)(2 * y) * ws + 2 * x) * 3;
      const uint8_t* p1 = p0 + (size_t)ws * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
    } else {
      const AxisTap tx = axis_tap(x, wd, ws, false), ty = axis_tap(y, hd, hs, true);
      const uint8_t* r0 = s + (size_t)ty.i0 * ws * 3;
      const uint8_t* r1 = s + (size_t)ty.i1 * ws * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[tx.i0 * 3 + c] * tx.c0 + r0[tx.i1 * 3 + c] * tx.c1;
        const int h1 = r1[tx.i0 * 3 + c] * tx.c0 + r1[tx.i1 * 3 + c] * tx.c1;
        const int o = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v[c] = min(max(o, 0), 255);
      }
    }
    // ToTensor: HWC uint8 -> CHW float32 / 255 (IEEE division, as torch's .div(255)); flips act on the tensor
    int ox = x, oy = y;
    if (flips) {
      const uint8_t fl = flips[img];
      if (fl & 1) ox = wd - 1 - x;
      if (fl & 2) oy = hd - 1 - y;
    }
    float* o = dst + (size_t)img * 3 * hd * wd + (size_t)oy * wd + ox;
    const size_t plane = (size_t)hd * wd;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[(size_t)(bgr ? 2 - c : c) * plane] = __fdiv_rn((float)v[c], 255.f);
  }
}

}  // namespace

extern "C" int adb_image_u8_to_f32(const uint8_t* src, int32_t n, int32_t hs, int32_t ws, int64_t src_image_stride, int32_t bgr,
                                   const uint8_t* flips, float* dst, int32_t hd, int32_t wd, void* stream) {
  ADB_REQUIRE(src && dst && n > 0 && hs > 0 && ws > 0 && hd > 0 && wd > 0, "adb_image_u8_to_f32: bad arguments");
  ADB_REQUIRE(src_image_stride >= (int64_t)hs * ws * 3, "adb_image_u8_to_f32: image stride %lld shorter than one %dx%d image", (long long)src_image_stride, hs, ws);
  const long long total = (long long)n * hd * wd;
  const int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  image_u8_to_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, n, hs, ws, src_image_stride, bgr, flips, dst, hd, wd);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}
