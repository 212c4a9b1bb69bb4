// train.cu — the HBM-bound kernels of the training step: batch-statistics BatchNorm forward/backward,
// activation backward, the image / guidance heads forward+backward, AttentionBlock backward, pooling backward.
// All stream NHWC bf16 maps with 16-byte accesses; per-channel reductions publish per-block partials that a finalize
// kernel folds in block order in fp64 (deterministic).
//
// Reference arithmetic replaced: what autograd runs for nn.BatchNorm2d(train) / ReLU / Tanh / Sigmoid / clamp / the
// AttentionBlock inside loss.backward() (training/train_dehazing.py:90-92) on the modules of models/dehazing/*.py.
#include "adb_ptx.cuh"
#include "adb_host.h"
#include <algorithm>
#include <math.h>

namespace {

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u[i]);
    f[2 * i] = __low2float(b);
    f[2 * i + 1] = __high2float(b);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  q.x = adb::pack_bf16x2(f[0], f[1]); q.y = adb::pack_bf16x2(f[2], f[3]);
  q.z = adb::pack_bf16x2(f[4], f[5]); q.w = adb::pack_bf16x2(f[6], f[7]);
  return q;
}
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// Streaming 16-byte loads as volatile asm: the front end otherwise sinks each load of an unrolled batch next to its first
// use (seen in SASS: loads issued pair by pair between the compute of the previous pair), which leaves too few bytes in
// flight per SM to cover HBM latency.  Volatile asm keeps the batch together in program order.
__device__ __forceinline__ uint4 ld_stream_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {      // coherent: the buffer may be written by this kernel (in place)
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == ADB_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == ADB_ACT_TANH) return 1.f - y * y;
  if (act == ADB_ACT_SIGMOID) return y * (1.f - y);
  if (act == ADB_ACT_SIGMOID2) return 0.5f * (1.f - y * y);
  return 1.f;
}
__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == ADB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == ADB_ACT_TANH) return tanhf(v);
  if (act == ADB_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  if (act == ADB_ACT_SIGMOID2) return 2.f / (1.f + __expf(-v)) - 1.f;
  return v;
}

inline int grid_for(long long work_items, int threads, int sm_count, int waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)sm_count * waves;
  return (int)std::max<long long>(1, std::min(blocks, cap));
}
inline int sm_count() {
  adbh::DeviceInfo di;
  return adbh::device_info(&di) == ADB_OK ? di.sm_count : 148;
}

// ---- streaming elementwise kernels over [pixels][c] maps (16-byte items, G = c/8 items per pixel).
// The grid's thread count S is a multiple of G, so a thread's items t, t+S, t+2S, ... all sit in the SAME channel group:
// the per-channel parameters are loaded once, no division runs inside the loop, and kEwU items per stream are in flight
// per thread before the first is consumed (2048 threads x 16 B per SM alone cannot cover HBM latency).
constexpr int kEwU = 4;
constexpr int kEwThreads = 256;
inline int ew_grid(long long pixels, int G, int sms, int waves = 8) {
  int a = G, b = kEwThreads;
  while (b) { const int t = a % b; a = b; b = t; }          // a = gcd(G, 256)
  const int m = G / a;                                       // grid must be a multiple of m
  const long long items = pixels * G;
  long long blocks = (items + (long long)kEwThreads * kEwU - 1) / ((long long)kEwThreads * kEwU);
  blocks = std::min<long long>(blocks, (long long)sms * waves);
  blocks = std::max<long long>(m, blocks / m * m);
  return (int)blocks;
}
struct EwIt { long long p, dp; int g; };
__device__ __forceinline__ EwIt ew_begin(int G) {
  const long long S = (long long)gridDim.x * blockDim.x;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  EwIt it;
  it.p = t / G; it.g = (int)(t - it.p * G); it.dp = S / G;
  return it;
}

constexpr int kRedThreads = 256;
inline int red_blocks(long long pixels, int c, int sms) {
  const int G = c / 8, PY = std::max(1, kRedThreads / G);
  // <= 4 blocks per SM: the streaming kernels hold 2-4 blocks of 256 threads per SM anyway, and the finalize kernels'
  // fold (a handful of CTAs walking every block's partials) costs time proportional to the block count
  return (int)std::max<long long>(1, std::min<long long>((pixels + PY * 8 - 1) / (PY * 8), (long long)sms * 4));
}

// ------------------------------------------------------------------ per-channel reductions over a map
// block = G x PY threads (G = c/8), thread (py, g) streams 8 channels down the pixels py, py + PY*gridDim, ...
// MODE 0: a = sum z, b = sum z^2                                    (BatchNorm batch statistics)
// MODE 1: g = dy * act'(y); a = sum g, b = sum g * xhat; writes g   (BatchNorm / bias backward; xhat = (z-mean)*rstd)
template <int MODE>
__global__ void __launch_bounds__(kRedThreads)
chan_reduce_kernel(const __nv_bfloat16* __restrict__ z, int pitch_z, const __nv_bfloat16* __restrict__ dy, int pitch_dy,
                   const __nv_bfloat16* __restrict__ y, int pitch_y, long long pixels, int c, int act,
                   const float* __restrict__ mean, const float* __restrict__ rstd, __nv_bfloat16* __restrict__ g_out,
                   int pitch_g, float* __restrict__ partials) {
  const int G = c / 8;
  const int PY = blockDim.x / G;
  const int g = threadIdx.x % G, py = threadIdx.x / G;
  extern __shared__ float sm[];   // [PY][2][c]
  float a[8], b[8], mu[8], rs[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { a[q] = 0.f; b[q] = 0.f; mu[q] = 0.f; rs[q] = 0.f; }
  if (py < PY) {
    if (MODE == 1 && z) { load8f(mean + g * 8, mu); load8f(rstd + g * 8, rs); }
    const long long step = (long long)gridDim.x * PY;
    for (long long p0 = (long long)blockIdx.x * PY + py; p0 < pixels; p0 += 2 * step) {
      const long long p1 = p0 + step;
      const bool two = p1 < pixels;
      if (MODE == 0) {
        const uint4 r0 = ld_stream_nc(z + (size_t)p0 * pitch_z + g * 8);
        const uint4 r1 = ld_stream_nc(z + (size_t)(two ? p1 : p0) * pitch_z + g * 8);
        float f[8], h[8];
        unpack8(r0, f); unpack8(r1, h);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (!two) h[q] = 0.f;
          a[q] += f[q] + h[q]; b[q] = fmaf(f[q], f[q], fmaf(h[q], h[q], b[q]));
        }
      } else {
        // both pixels' loads are issued before either is consumed
        uint4 rd[2], ry[2], rz[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const long long p = (u && two) ? p1 : p0;      // a dead second slot re-reads p0 (never consumed): unconditional loads
          rd[u] = ld_stream_nc(dy + (size_t)p * pitch_dy + g * 8);
          ry[u] = y ? ld_stream_nc(y + (size_t)p * pitch_y + g * 8) : make_uint4(0, 0, 0, 0);
          rz[u] = z ? ld_stream_nc(z + (size_t)p * pitch_z + g * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !two) break;
          const long long p = u ? p1 : p0;
          float d[8], yy[8], zz[8];
          unpack8(rd[u], d); unpack8(ry[u], yy); unpack8(rz[u], zz);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float gg = y ? d[q] * act_grad(yy[q], act) : d[q];
            d[q] = gg;
            a[q] += gg;
            if (z) b[q] = fmaf(gg, (zz[q] - mu[q]) * rs[q], b[q]);
          }
          if (g_out) *reinterpret_cast<uint4*>(g_out + (size_t)p * pitch_g + g * 8) = pack8(d);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sm[(py * 2 + 0) * c + g * 8 + q] = a[q];
      sm[(py * 2 + 1) * c + g * 8 + q] = b[q];
    }
  }
  __syncthreads();
  float* mine = partials + (size_t)blockIdx.x * 2 * c;
  for (int ch = threadIdx.x; ch < 2 * c; ch += blockDim.x) {
    const int which = ch / c, cc = ch - which * c;
    float s = 0.f;
    for (int r = 0; r < PY; ++r) s += sm[(r * 2 + which) * c + cc];
    mine[ch] = s;
  }
}

// Fold the per-block partials [nblocks][2][c] of 8 channels per CTA: 32 lanes per channel each sum every 32nd block in
// fp64, then one thread per (channel, which) adds the 32 lane sums in lane order — a fixed tree, so results are
// deterministic.  Returns (through shared memory) sums[which][8].  blockDim = 256 = 32 lanes x 8 channels.
constexpr int kFoldLanes = 32;
__device__ __forceinline__ void fold_partials8(const float* __restrict__ partials, int nblocks, int c, int ch0, double (*s_out)[8]) {
  __shared__ double s_lane[2][kFoldLanes][8];
  const int chl = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int ch = ch0 + chl;
  double a = 0.0, b = 0.0;
  if (ch < c) {
    // eight blocks' partials are loaded before any is added (the loads are independent; the fp64 adds stay in block order)
    int k = lane;
    for (; k + 7 * kFoldLanes < nblocks; k += 8 * kFoldLanes) {
      float va[8], vb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        va[u] = __ldg(partials + (size_t)(k + u * kFoldLanes) * 2 * c + ch);
        vb[u] = __ldg(partials + (size_t)(k + u * kFoldLanes) * 2 * c + c + ch);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { a += (double)va[u]; b += (double)vb[u]; }
    }
    for (; k < nblocks; k += kFoldLanes) {
      a += (double)partials[(size_t)k * 2 * c + ch];
      b += (double)partials[(size_t)k * 2 * c + c + ch];
    }
  }
  s_lane[0][lane][chl] = a;
  s_lane[1][lane][chl] = b;
  __syncthreads();
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, cc = threadIdx.x & 7;
    double t = 0.0;
    for (int l = 0; l < kFoldLanes; ++l) t += s_lane[which][l][cc];
    s_out[which][cc] = t;
  }
  __syncthreads();
}

// First fold of the conv epilogue's channel partials (adb_conv_desc.stat_out: [slots][2][cpitch], one slot per 32 output pixels):
// CTA (x, y) sums slot chunk y for channels 32x..32x+31 in fp64, slot order fixed, and writes [y][2][c] float partials for
// bn_finalize_kernel.  block = 32 channels x 8 slot lanes: a warp reads 128 contiguous bytes of one slot.
__global__ void stat_fold_kernel(const float* __restrict__ stat, long long slots, int cpitch, int c, float* __restrict__ out) {
  __shared__ double s_part[2][8][32];
  const int ch = blockIdx.x * 32 + threadIdx.x, ly = threadIdx.y;
  const long long per = (slots + gridDim.y - 1) / gridDim.y;
  const long long k0 = (long long)blockIdx.y * per, k1 = min(slots, k0 + per);
  double a = 0.0, b = 0.0;
  if (ch < c) {
    long long k = k0 + ly;
    for (; k + 24 < k1; k += 32) {         // four slots in flight per thread
      float va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        va[u] = __ldg(stat + (size_t)(k + 8 * u) * 2 * cpitch + ch);
        vb[u] = __ldg(stat + (size_t)(k + 8 * u) * 2 * cpitch + cpitch + ch);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a += (double)va[u]; b += (double)vb[u]; }
    }
    for (; k < k1; k += 8) {
      a += (double)__ldg(stat + (size_t)k * 2 * cpitch + ch);
      b += (double)__ldg(stat + (size_t)k * 2 * cpitch + cpitch + ch);
    }
  }
  s_part[0][ly][threadIdx.x] = a;
  s_part[1][ly][threadIdx.x] = b;
  __syncthreads();
  if (ly < 2 && ch < c) {
    double t = 0.0;
    for (int l = 0; l < 8; ++l) t += s_part[ly][l][threadIdx.x];
    out[((size_t)blockIdx.y * 2 + ly) * c + ch] = (float)t;
  }
}

// BatchNorm2d(train) statistics -> the epilogue affine, plus the running-statistics update (momentum, unbiased variance).
// grid = c/8 CTAs of 256 threads.
__global__ void bn_finalize_kernel(const float* __restrict__ partials, int nblocks, int c, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                   float* running_mean, float* running_var, long long* num_batches_tracked,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  __shared__ double s_sum[2][8];
  fold_partials8(partials, nblocks, c, blockIdx.x * 8, s_sum);
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (threadIdx.x >= 8) return;
  const int ch = blockIdx.x * 8 + threadIdx.x;
  if (ch >= c) return;
  const double s = s_sum[0][threadIdx.x], ss = s_sum[1][threadIdx.x];
  const double m = s / count;
  double var = ss / count - m * m;
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  mean[ch] = (float)m;
  rstd[ch] = r;
  const float sc = gamma[ch] * r;
  scale[ch] = sc;
  shift[ch] = beta[ch] - (float)m * sc;
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// y = act(z*scale + shift (+ residual))
__global__ void __launch_bounds__(kEwThreads)
affine_act_kernel(const __nv_bfloat16* __restrict__ z, int pitch_z, long long pixels, int c,
                  const float* __restrict__ scale, const float* __restrict__ shift,
                  const __nv_bfloat16* __restrict__ res, int pitch_r, int act, __nv_bfloat16* __restrict__ y,
                  int pitch_y) {
  const EwIt it = ew_begin(c / 8);
  const int co = it.g * 8;
  float sc[8], sh[8];
  load8f(scale + co, sc); load8f(shift + co, sh);
  for (long long p0 = it.p; p0 < pixels; p0 += kEwU * it.dp) {
    uint4 rz[kEwU], rr[kEwU];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp < pixels ? p0 + u * it.dp : p0;     // dead slots re-read p0: loads stay unconditional
      rz[u] = ld_stream_nc(z + (size_t)p * pitch_z + co);
      if (res) rr[u] = ld_stream_nc(res + (size_t)p * pitch_r + co);
    }
    // (pixel-outer compute here: the channel-outer form costs 128 registers with the activation switch and measured slower)
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp;
      if (p >= pixels) break;
      float f[8], r[8];
      unpack8(rz[u], f);
      if (res) unpack8(rr[u], r);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = act_fwd(fmaf(f[q], sc[q], sh[q]) + (res ? r[q] : 0.f), act);
      *reinterpret_cast<uint4*>(y + (size_t)p * pitch_y + co) = pack8(f);
    }
  }
}

// BatchNorm backward coefficients: dz = A*g + B*z + C per channel; dgamma = sum g*xhat, dbeta = sum g.
// gamma == NULL: bias-only layer -> only dbeta (the conv bias gradient) is produced.
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, int c, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, int accumulate, float* dgamma, float* dbeta,
                                       float* __restrict__ coef) {
  __shared__ double s_sum[2][8];
  fold_partials8(partials, nblocks, c, blockIdx.x * 8, s_sum);
  if (threadIdx.x >= 8) return;
  const int ch = blockIdx.x * 8 + threadIdx.x;
  if (ch >= c) return;
  const double sg = s_sum[0][threadIdx.x], sgx = s_sum[1][threadIdx.x];
  if (dbeta) dbeta[ch] = (accumulate ? dbeta[ch] : 0.f) + (float)sg;
  if (!gamma) return;
  if (dgamma) dgamma[ch] = (accumulate ? dgamma[ch] : 0.f) + (float)sgx;
  const double k = (double)gamma[ch] * (double)rstd[ch];
  const double mg = sg / count, mgx = sgx / count;
  const double B = -k * mgx * (double)rstd[ch];
  coef[ch] = (float)k;
  coef[c + ch] = (float)B;
  coef[2 * c + ch] = (float)(-k * mg - B * (double)mean[ch]);
}

__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* g, int pitch_g, const __nv_bfloat16* __restrict__ z,
                    int pitch_z, long long pixels, int c, const float* __restrict__ coef,
                    __nv_bfloat16* dz, int pitch_dz) {
  const EwIt it = ew_begin(c / 8);
  const int co = it.g * 8;
  float A[8], B[8], Cc[8];
  load8f(coef + co, A); load8f(coef + c + co, B); load8f(coef + 2 * c + co, Cc);
  for (long long p0 = it.p; p0 < pixels; p0 += kEwU * it.dp) {
    uint4 rg[kEwU], rz[kEwU];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp < pixels ? p0 + u * it.dp : p0;
      rg[u] = ld_stream(g + (size_t)p * pitch_g + co);        // dz may alias g: plain loads
      rz[u] = ld_stream_nc(z + (size_t)p * pitch_z + co);
    }
    float gg[kEwU][8], zz[kEwU][8];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) { unpack8(rg[u], gg[u]); unpack8(rz[u], zz[u]); }
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int u = 0; u < kEwU; ++u) gg[u][q] = fmaf(A[q], gg[u][q], fmaf(B[q], zz[u][q], Cc[q]));
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp;
      if (p < pixels) *reinterpret_cast<uint4*>(dz + (size_t)p * pitch_dz + co) = pack8(gg[u]);
    }
  }
}

// ---- BatchNorm + ReLU backward without the activation output: the ReLU mask is recomputed from z with the forward's own
// (scale, shift) — fmaf(z, scale, shift) > 0 is exactly the sign test affine_act_kernel's output passed — so neither y is
// read nor g = dy*mask written and re-read: 2 + 3 streamed passes instead of 4 + 3 (no residual input: y = relu(BN(z))).
__global__ void __launch_bounds__(kRedThreads)
bn_relu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ z, int pitch_z, const __nv_bfloat16* __restrict__ dy, int pitch_dy,
                          long long pixels, int c, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ partials) {
  const int G = c / 8;
  const int PY = blockDim.x / G;
  const int g = threadIdx.x % G, py = threadIdx.x / G;
  extern __shared__ float sm[];   // [PY][2][c]
  float a[8], b[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { a[q] = 0.f; b[q] = 0.f; }
  if (py < PY) {
    float mu[8], rs[8], sc[8], sh[8];
    load8f(mean + g * 8, mu); load8f(rstd + g * 8, rs); load8f(scale + g * 8, sc); load8f(shift + g * 8, sh);
    const long long step = (long long)gridDim.x * PY;
    for (long long p0 = (long long)blockIdx.x * PY + py; p0 < pixels; p0 += 4 * step) {
      uint4 rd[4], rz[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // unconditional loads (a dead slot re-reads p0) so that all eight are issued before the first is consumed
        const long long p = p0 + u * step < pixels ? p0 + u * step : p0;
        rd[u] = ld_stream_nc(dy + (size_t)p * pitch_dy + g * 8);
        rz[u] = ld_stream_nc(z + (size_t)p * pitch_z + g * 8);
      }
      float d[4][8], zz[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) { unpack8(rd[u], d[u]); unpack8(rz[u], zz[u]); }
      // channel-outer / pixel-inner: the first channel's sums need all eight loads, so none can be sunk below its use
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool live = p0 + u * step < pixels;
          const float gg = (live && fmaf(zz[u][q], sc[q], sh[q]) > 0.f) ? d[u][q] : 0.f;
          sa += gg;
          sb = fmaf(gg, (zz[u][q] - mu[q]) * rs[q], sb);
        }
        a[q] += sa;
        b[q] += sb;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sm[(py * 2 + 0) * c + g * 8 + q] = a[q];
      sm[(py * 2 + 1) * c + g * 8 + q] = b[q];
    }
  }
  __syncthreads();
  float* mine = partials + (size_t)blockIdx.x * 2 * c;
  for (int ch = threadIdx.x; ch < 2 * c; ch += blockDim.x) {
    const int which = ch / c, cc = ch - which * c;
    float s = 0.f;
    for (int r = 0; r < PY; ++r) s += sm[(r * 2 + which) * c + cc];
    mine[ch] = s;
  }
}

// dz (+)= A*g + B*z + C with g = dy * [fmaf(z, scale, shift) > 0]; dz may alias dy.  ACC: dz is a gradient buffer that
// already holds other consumers' contributions (DenseNet block buffer prefix) and is read-modified-written.
template <bool ACC>
__global__ void __launch_bounds__(kEwThreads)
bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int pitch_dy, const __nv_bfloat16* __restrict__ z,
                         int pitch_z, long long pixels, int c, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ coef,
                         __nv_bfloat16* dz, int pitch_dz) {
  const EwIt it = ew_begin(c / 8);
  const int co = it.g * 8;
  float A[8], B[8], Cc[8], sc[8], sh[8];
  load8f(coef + co, A); load8f(coef + c + co, B); load8f(coef + 2 * c + co, Cc);
  load8f(scale + co, sc); load8f(shift + co, sh);
  for (long long p0 = it.p; p0 < pixels; p0 += kEwU * it.dp) {
    uint4 rd[kEwU], rz[kEwU], rp[kEwU];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp < pixels ? p0 + u * it.dp : p0;
      rd[u] = ld_stream(dy + (size_t)p * pitch_dy + co);     // dz may alias dy: plain loads
      rz[u] = ld_stream_nc(z + (size_t)p * pitch_z + co);
      if (ACC) rp[u] = ld_stream(dz + (size_t)p * pitch_dz + co);
    }
    float gg[kEwU][8], zz[kEwU][8], prev[kEwU][8];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      unpack8(rd[u], gg[u]); unpack8(rz[u], zz[u]);
      if (ACC) unpack8(rp[u], prev[u]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int u = 0; u < kEwU; ++u) {
        const float g = fmaf(zz[u][q], sc[q], sh[q]) > 0.f ? gg[u][q] : 0.f;
        float v = fmaf(A[q], g, fmaf(B[q], zz[u][q], Cc[q]));
        if (ACC) v = prev[u][q] + __bfloat162float(__float2bfloat16(v));     // same rounding as a separate dz map + adb_add_bf16
        gg[u][q] = v;
      }
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp;
      if (p < pixels) *reinterpret_cast<uint4*>(dz + (size_t)p * pitch_dz + co) = pack8(gg[u]);
    }
  }
}

__global__ void __launch_bounds__(kEwThreads)
add_bf16_kernel(__nv_bfloat16* __restrict__ a, int pitch_a, const __nv_bfloat16* __restrict__ b, int pitch_b,
                long long pixels, int c) {
  const EwIt it = ew_begin(c / 8);
  const int co = it.g * 8;
  for (long long p0 = it.p; p0 < pixels; p0 += kEwU * it.dp) {
    uint4 ra[kEwU], rb[kEwU];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp < pixels ? p0 + u * it.dp : p0;
      ra[u] = ld_stream(a + (size_t)p * pitch_a + co);
      rb[u] = ld_stream_nc(b + (size_t)p * pitch_b + co);
    }
    float x[kEwU][8], y[kEwU][8];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) { unpack8(ra[u], x[u]); unpack8(rb[u], y[u]); }
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int u = 0; u < kEwU; ++u) x[u][q] += y[u][q];
#pragma unroll
    for (int u = 0; u < kEwU; ++u) {
      const long long p = p0 + u * it.dp;
      if (p < pixels) *reinterpret_cast<uint4*>(a + (size_t)p * pitch_a + co) = pack8(x[u]);
    }
  }
}

// ------------------------------------------------------------------ image head (low:45, medium:117, high:135-138)
// v = act(z[..., c]) per colour; BLEND: out = (1-alpha) x + alpha v; RESIDUAL: clamp(x + v); GUIDED: clamp(x + v*guidance)
__global__ void img_head_fwd_kernel(const __nv_bfloat16* __restrict__ z, int pitch, const float* __restrict__ x,
                                    const float* __restrict__ guidance, const float* __restrict__ alpha_p, int mode, int act,
                                    int n, long long hw, float* __restrict__ out) {
  const long long total = (long long)n * hw;
  const float alpha = mode == ADB_IMG_BLEND ? __ldg(alpha_p) : 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw, q = p - img * hw;
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(z + (size_t)p * pitch));
    const __nv_bfloat162 b01 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b23 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const float zz[3] = {__low2float(b01), __high2float(b01), __low2float(b23)};
    const float gd = mode == ADB_IMG_GUIDED ? __ldg(guidance + p) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)img * 3 + c) * hw + q;
      const float xv = __ldg(x + o), v = act_fwd(zz[c], act);
      out[o] = mode == ADB_IMG_BLEND ? (1.f - alpha) * xv + alpha * v : fminf(fmaxf(xv + v * gd, 0.f), 1.f);
    }
  }
}

// red[0..2] += dbias, red[3] += dalpha  (block partial + one atomic per block and value)
__global__ void img_head_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ z, int pitch,
                                    const float* __restrict__ x, const float* __restrict__ guidance,
                                    const float* __restrict__ alpha_p, int mode, int act, int n, long long hw,
                                    __nv_bfloat16* __restrict__ dz, float* __restrict__ dguidance, float* __restrict__ red) {
  const long long total = (long long)n * hw;
  const float alpha = mode == ADB_IMG_BLEND ? __ldg(alpha_p) : 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw, q = p - img * hw;
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(z + (size_t)p * pitch));
    const __nv_bfloat162 b01 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b23 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const float zz[3] = {__low2float(b01), __high2float(b01), __low2float(b23)};
    const float gd = mode == ADB_IMG_GUIDED ? __ldg(guidance + p) : 1.f;
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dgd = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)img * 3 + c) * hw + q;
      const float xv = __ldg(x + o), v = act_fwd(zz[c], act), go = __ldg(dout + o);
      float dv;
      if (mode == ADB_IMG_BLEND) {
        dv = alpha * go;
        acc[3] += go * (v - xv);
      } else {
        const float pre = xv + v * gd;
        const float m = (pre >= 0.f && pre <= 1.f) ? go : 0.f;   // torch.clamp passes the gradient on the closed interval
        dv = m * gd;
        dgd += m * v;
      }
      d[c] = dv * act_grad(v, act);
      acc[c] += d[c];
    }
    uint4* o16 = reinterpret_cast<uint4*>(dz + (size_t)p * pitch);
    o16[0] = pack8(d);
    for (int k = 1; k < pitch / 8; ++k) o16[k] = make_uint4(0, 0, 0, 0);
    if (dguidance) dguidance[p] = dgd;
  }
  __shared__ float s_red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s_red[threadIdx.x][w];
    atomicAdd(red + threadIdx.x, v);
  }
}

// ------------------------------------------------------------------ 1x1 sigmoid head (detail_branch tail, high:87-89)
// g = sigmoid(dot(y[..., :c], w) + b) ; backward: dpre = dg * g(1-g); dy = dpre * w; red[0..c) += dpre*y, red[c] += dpre
__global__ void dot_head_fwd_kernel(const __nv_bfloat16* __restrict__ y, int pitch, int c, const float* __restrict__ w,
                                    const float* __restrict__ b, long long pixels, float* __restrict__ g) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    float acc = __ldg(b);
    for (int k = 0; k < c / 8; ++k) {
      float f[8], ww[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(y + (size_t)p * pitch + k * 8)), f);
      load8f(w + k * 8, ww);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(f[q], ww[q], acc);
    }
    g[p] = 1.f / (1.f + __expf(-acc));
  }
}

template <int C>
__global__ void dot_head_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                    int pitch, const float* __restrict__ w, long long pixels, __nv_bfloat16* __restrict__ dy,
                                    int pitch_dy, float* __restrict__ red) {
  float acc[C + 1];
#pragma unroll
  for (int i = 0; i <= C; ++i) acc[i] = 0.f;
  float ww[C];
#pragma unroll
  for (int i = 0; i < C; ++i) ww[i] = __ldg(w + i);
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    const float gv = __ldg(g + p);
    const float dpre = __ldg(dg + p) * gv * (1.f - gv);
    acc[C] += dpre;
#pragma unroll
    for (int k = 0; k < C / 8; ++k) {
      float f[8], d[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(y + (size_t)p * pitch + k * 8)), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) { acc[k * 8 + q] = fmaf(dpre, f[q], acc[k * 8 + q]); d[q] = dpre * ww[k * 8 + q]; }
      *reinterpret_cast<uint4*>(dy + (size_t)p * pitch_dy + k * 8) = pack8(d);
    }
  }
  __shared__ float s_red[C + 1][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= C; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x <= C) {
    float v = 0.f;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) v += s_red[threadIdx.x][wi];
    atomicAdd(red + threadIdx.x, v);
  }
}

// ------------------------------------------------------------------ AttentionBlock backward (base_model.py:64-78)
// forward: cg = sigmoid(fc(avg)+fc(max)) [n,c]; u = x*cg; stats = (mean_c u, max_c u); sg = sigmoid(conv7x7(stats)); y = u*sg
// pass 1 (per pixel): d_pre[p] = (sum_c dy*u) * sg(1-sg);  amax[p] = first argmax_c u
template <int LP, int ML>
__global__ void attn_bwd_pixel_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, int n,
                                      long long hw, int c, const float* __restrict__ gate, const float* __restrict__ sg,
                                      float* __restrict__ d_pre, int* __restrict__ amax) {
  const int G = c / 8;
  const int sub = threadIdx.x % LP;
  constexpr int PPW = 32 / LP;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long stride = ((long long)gridDim.x * blockDim.x >> 5) * PPW;
  const long long total = (long long)n * hw;
  for (long long pw = warp_id * PPW; pw < total; pw += stride) {
    const long long p = pw + (threadIdx.x & 31) / LP;
    const bool live = p < total;
    float s = 0.f, m = -INFINITY;
    int mi = 0x7fffffff;
    if (live) {
      const int img = (int)(p / hw);
      const float* gt = gate + (size_t)img * c;
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        const int g = sub + k * LP;
        if (g < G) {
          float f[8], d[8], gg[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(x + (size_t)p * c + g * 8)), f);
          unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (size_t)p * c + g * 8)), d);
          load8f(gt + g * 8, gg);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float u = f[q] * gg[q];
            s = fmaf(d[q], u, s);
            if (u > m) { m = u; mi = g * 8 + q; }     // ascending channel order within the lane: first max wins
          }
        }
      }
    }
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    if (live && sub == 0) {
      const float sv = __ldg(sg + p);
      d_pre[p] = s * sv * (1.f - sv);
      amax[p] = mi;
    }
  }
}

// pass 2 (stencil transpose): d_stats[ch][q] = sum_off wsp[ch][off] * d_pre[q - off];  dwsp[ch][off] += sum_p d_pre[p] * stats[ch][p + off]
constexpr int kBwTW = 32, kBwTH = 16;
__global__ void attn_bwd_stencil_kernel(const float* __restrict__ d_pre, const float* __restrict__ stats, int h, int w,
                                        const float* __restrict__ wsp, float* __restrict__ d_stats, float* __restrict__ dwsp) {
  __shared__ float s_w[98];
  __shared__ float s_d[kBwTH + 6][kBwTW + 6];
  __shared__ float2 s_t[kBwTH + 6][kBwTW + 6];
  __shared__ float s_acc[98];
  const int img = blockIdx.z;
  const int tid = threadIdx.y * kBwTW + threadIdx.x;
  for (int i = tid; i < 98; i += kBwTW * kBwTH) { s_w[i] = wsp[i]; s_acc[i] = 0.f; }
  const int x0 = blockIdx.x * kBwTW - 3, y0 = blockIdx.y * kBwTH - 3;
  const float* dp = d_pre + (size_t)img * h * w;
  const float2* st = reinterpret_cast<const float2*>(stats) + (size_t)img * h * w;
  for (int i = tid; i < (kBwTH + 6) * (kBwTW + 6); i += kBwTW * kBwTH) {
    const int ty = i / (kBwTW + 6), tx = i - ty * (kBwTW + 6);
    const int yy = y0 + ty, xx = x0 + tx;
    const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
    s_d[ty][tx] = in ? __ldg(dp + (size_t)yy * w + xx) : 0.f;
    s_t[ty][tx] = in ? __ldg(st + (size_t)yy * w + xx) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int px = blockIdx.x * kBwTW + threadIdx.x, py = blockIdx.y * kBwTH + threadIdx.y;
  const bool live = px < w && py < h;
  // d_stats at q = (py, px): forward pre[p] = sum_off w[off] * stats[p + off - 3]  =>  d_stats[q] = sum_off w[off] * d_pre[q - off + 3]
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy)
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      const float v = s_d[threadIdx.y + 6 - dy][threadIdx.x + 6 - dx];
      a0 = fmaf(s_w[dy * 7 + dx], v, a0);
      a1 = fmaf(s_w[49 + dy * 7 + dx], v, a1);
    }
  if (live) *reinterpret_cast<float2*>(d_stats + (((size_t)img * h + py) * w + px) * 2) = make_float2(a0, a1);
  // dwsp: this thread's pixel p contributes d_pre[p] * stats[p + off - 3]
  const float dcen = live ? s_d[threadIdx.y + 3][threadIdx.x + 3] : 0.f;
  const int lane = tid & 31;
  for (int off = 0; off < 49; ++off) {
    const int dy = off / 7, dx = off - dy * 7;
    const float2 v = s_t[threadIdx.y + dy][threadIdx.x + dx];
    float c0 = dcen * v.x, c1 = dcen * v.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); }
    if (lane == 0) { atomicAdd(&s_acc[off], c0); atomicAdd(&s_acc[49 + off], c1); }
  }
  __syncthreads();
  for (int i = tid; i < 98; i += kBwTW * kBwTH) atomicAdd(dwsp + i, s_acc[i]);
}

// pass 3: du = dy*sg + d_mean/C + d_max*[c == amax];  dx = du*cg (bf16);  per-block partial of dgate[n][c] = sum_p du*x;
//         pos[n][c] = min pixel index with x == max (AdaptiveMaxPool2d's winner), for pass 5.
// grid = (chunks, n); block = G x PY
__global__ void __launch_bounds__(kRedThreads)
attn_bwd_main_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, long long hw, int c,
                     const float* __restrict__ gate, const float* __restrict__ sg, const float* __restrict__ d_stats,
                     const int* __restrict__ amax, const float* __restrict__ pool, __nv_bfloat16* __restrict__ dx,
                     float* __restrict__ partials, int* __restrict__ pos) {
  const int img = blockIdx.y, chunks = gridDim.x;
  const int G = c / 8;
  const int PY = blockDim.x / G;
  const int g = threadIdx.x % G, py = threadIdx.x / G;
  extern __shared__ float sm[];   // [PY][c]
  float a[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a[q] = 0.f;
  if (py < PY) {
    float gg[8], mx[8];
    load8f(gate + (size_t)img * c + g * 8, gg);
    load8f(pool + ((size_t)img * 2 + 1) * c + g * 8, mx);
    const float inv_c = 1.f / (float)c;
    int best[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) best[q] = 0x7fffffff;
    for (long long p = (long long)blockIdx.x * PY + py; p < hw; p += (long long)chunks * PY) {
      const size_t gp = (size_t)img * hw + p;
      float f[8], d[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + gp * c + g * 8)), f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + gp * c + g * 8)), d);
      const float sv = __ldg(sg + gp);
      const float2 ds = __ldg(reinterpret_cast<const float2*>(d_stats) + gp);
      const int am = __ldg(amax + gp);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float du = fmaf(d[q], sv, ds.x * inv_c);
        if (am == g * 8 + q) du += ds.y;
        a[q] = fmaf(du, f[q], a[q]);
        d[q] = du * gg[q];
        if (f[q] == mx[q] && (int)p < best[q]) best[q] = (int)p;
      }
      *reinterpret_cast<uint4*>(dx + gp * c + g * 8) = pack8(d);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sm[py * c + g * 8 + q] = a[q];
      if (best[q] != 0x7fffffff) atomicMin(pos + (size_t)img * c + g * 8 + q, best[q]);
    }
  }
  __syncthreads();
  float* mine = partials + ((size_t)img * chunks + blockIdx.x) * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < PY; ++r) s += sm[r * c + ch];
    mine[ch] = s;
  }
}

// pass 4 (one block per image): dgate -> through sigmoid and the shared MLP (fc = W2 relu(W1 .)) for the avg and max inputs.
// d_avg[n][c] (already divided by hw), d_max[n][c]; dw1[cr][c], dw2[c][cr] accumulate over images with atomics.
__global__ void attn_bwd_gate_kernel(const float* __restrict__ partials, int chunks, const float* __restrict__ pool, float inv_hw,
                                     int c, int cr, const float* __restrict__ gate, const float* __restrict__ w1,
                                     const float* __restrict__ w2, float* __restrict__ d_avg, float* __restrict__ d_max,
                                     float* __restrict__ dw1, float* __restrict__ dw2) {
  const int img = blockIdx.x;
  extern __shared__ float sm[];   // avg[c], mx[c], dpre[c], ha[cr], hm[cr], dha[cr], dhm[cr]
  float* avg = sm; float* mx = sm + c; float* dpre = sm + 2 * c;
  float* ha = sm + 3 * c; float* hm = ha + cr; float* dha = hm + cr; float* dhm = dha + cr;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    avg[ch] = pool[((size_t)img * 2 + 0) * c + ch] * inv_hw;
    mx[ch] = pool[((size_t)img * 2 + 1) * c + ch];
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += partials[((size_t)img * chunks + k) * c + ch];
    const float gv = gate[(size_t)img * c + ch];
    dpre[ch] = s * gv * (1.f - gv);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < cr; j += nwarps) {          // hidden pre-activations and their gradients
    float a = 0.f, b = 0.f, dh = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      const float wv = w1[(size_t)j * c + ch];
      a = fmaf(wv, avg[ch], a); b = fmaf(wv, mx[ch], b);
      dh = fmaf(w2[(size_t)ch * cr + j], dpre[ch], dh);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); dh += __shfl_xor_sync(0xffffffffu, dh, o);
    }
    if (lane == 0) { ha[j] = fmaxf(a, 0.f); hm[j] = fmaxf(b, 0.f); dha[j] = a > 0.f ? dh : 0.f; dhm[j] = b > 0.f ? dh : 0.f; }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float da = 0.f, dm = 0.f;
    for (int j = 0; j < cr; ++j) {
      const float wv = w1[(size_t)j * c + ch];
      da = fmaf(wv, dha[j], da); dm = fmaf(wv, dhm[j], dm);
      atomicAdd(dw1 + (size_t)j * c + ch, dha[j] * avg[ch] + dhm[j] * mx[ch]);
      atomicAdd(dw2 + (size_t)ch * cr + j, dpre[ch] * (ha[j] + hm[j]));
    }
    d_avg[(size_t)img * c + ch] = da * inv_hw;
    d_max[(size_t)img * c + ch] = dm;
  }
}

// pass 5: dx[p][c] += d_avg[n][c] + d_max[n][c] * [p == pos[n][c]]
__global__ void attn_bwd_pool_kernel(__nv_bfloat16* __restrict__ dx, int n, long long hw, int c, const float* __restrict__ d_avg,
                                     const float* __restrict__ d_max, const int* __restrict__ pos) {
  const int G = c / 8;
  const long long total = (long long)n * hw * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long gp = t / G;
    const int img = (int)(gp / hw);
    const int p = (int)(gp - (long long)img * hw);
    float f[8], da[8], dm[8];
    unpack8(*reinterpret_cast<const uint4*>(dx + (size_t)gp * c + g * 8), f);
    load8f(d_avg + (size_t)img * c + g * 8, da);
    load8f(d_max + (size_t)img * c + g * 8, dm);
    const int4 p0 = __ldg(reinterpret_cast<const int4*>(pos + (size_t)img * c + g * 8));
    const int4 p1 = __ldg(reinterpret_cast<const int4*>(pos + (size_t)img * c + g * 8 + 4));
    const int pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] += da[q] + (pp[q] == p ? dm[q] : 0.f);
    *reinterpret_cast<uint4*>(dx + (size_t)gp * c + g * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------ soft/gated blend backward (routing.py:111-127)
// out = sum_k w[b][k] y_k :  dy_k = w[b][k] * dout ;  dwt[b][k] = sum_chw dout * y_k  (block partial + atomics)
__global__ void blend3_bwd_kernel(const float4* __restrict__ dout, const float4* __restrict__ y0, const float4* __restrict__ y1,
                                  const float4* __restrict__ y2, const float* __restrict__ wts, long long chw4,
                                  float4* __restrict__ d0, float4* __restrict__ d1, float4* __restrict__ d2, float* __restrict__ dwt) {
  const int b = blockIdx.y;
  const float w0 = __ldg(wts + b * 3), w1 = __ldg(wts + b * 3 + 1), w2 = __ldg(wts + b * 3 + 2);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const size_t base = (size_t)b * chw4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < chw4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = __ldg(dout + base + i);
    const float4 p = __ldg(y0 + base + i), q = __ldg(y1 + base + i), r = __ldg(y2 + base + i);
    a0 += g.x * p.x + g.y * p.y + g.z * p.z + g.w * p.w;
    a1 += g.x * q.x + g.y * q.y + g.z * q.z + g.w * q.w;
    a2 += g.x * r.x + g.y * r.y + g.z * r.z + g.w * r.w;
    d0[base + i] = make_float4(w0 * g.x, w0 * g.y, w0 * g.z, w0 * g.w);
    d1[base + i] = make_float4(w1 * g.x, w1 * g.y, w1 * g.z, w1 * g.w);
    d2[base + i] = make_float4(w2 * g.x, w2 * g.y, w2 * g.z, w2 * g.w);
  }
  __shared__ float s_red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float v[3] = {a0, a1, a2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) s_red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += s_red[threadIdx.x][wi];
    atomicAdd(dwt + b * 3 + threadIdx.x, t);
  }
}
// softmax(l / T) backward: dl_k = w_k (dw_k - sum_j w_j dw_j) / T
__global__ void softmax3_bwd_kernel(const float* __restrict__ wts, const float* __restrict__ dwt, float inv_t, int b,
                                    float* __restrict__ dlogits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const float w0 = wts[i * 3], w1 = wts[i * 3 + 1], w2 = wts[i * 3 + 2];
  const float g0 = dwt[i * 3], g1 = dwt[i * 3 + 1], g2 = dwt[i * 3 + 2];
  const float dot = w0 * g0 + w1 * g1 + w2 * g2;
  dlogits[i * 3] = w0 * (g0 - dot) * inv_t;
  dlogits[i * 3 + 1] = w1 * (g1 - dot) * inv_t;
  dlogits[i * 3 + 2] = w2 * (g2 - dot) * inv_t;
}

// ------------------------------------------------------------------ perceptual losses (loss.py:47-108)
// per-colour affine of an NCHW fp32 image batch: y = x*scale[c] + shift[c]  (ImageNet / LPIPS input normalisation)
__global__ void image_affine_kernel(const float* __restrict__ x, long long hw, long long total, float s0, float s1, float s2,
                                    float b0, float b1, float b2, float* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / hw) % 3);
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), sh = c == 0 ? b0 : (c == 1 ? b1 : b2);
    y[i] = fmaf(x[i], sc, sh);
  }
}

// nn.MaxPool2d(k, stride, pad) on NHWC bf16 (VGG: 2,2,0; AlexNet: 3,2,0; ResNet stem: 3,2,1), forward and backward.
// Backward is a gather: every input pixel visits the (<= 4) windows that contain it and takes dy where it is the
// window's FIRST maximum in scan order (ties resolved like the reference's CPU kernel; never double counted).
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, int k, int stride, int pad,
                                   int ho, int wo, __nv_bfloat16* __restrict__ y) {
  const int G = c / 8;
  const long long total = (long long)n * ho * wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int img = (int)(p / ho);
    float m[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) m[q] = -INFINITY;
    for (int r = 0; r < k; ++r) {
      const int yy = yo * stride - pad + r;
      if (yy < 0 || yy >= h) continue;
      for (int s = 0; s < k; ++s) {
        const int xx = xo * stride - pad + s;
        if (xx < 0 || xx >= w) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)img * h + yy) * w + xx) * c + g * 8)), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], f[q]);
      }
    }
    *reinterpret_cast<uint4*>(y + (((size_t)img * ho + yo) * wo + xo) * c + g * 8) = pack8(m);
  }
}

__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                   const __nv_bfloat16* __restrict__ y, int n, int h, int w, int c, int k, int stride, int pad,
                                   int ho, int wo, __nv_bfloat16* __restrict__ dx) {
  const int G = c / 8;
  const long long total = (long long)n * h * w * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xi = (int)(p % w); p /= w;
    const int yi = (int)(p % h);
    const int img = (int)(p / h);
    float me[8], acc[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)img * h + yi) * w + xi) * c + g * 8)), me);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    // windows (yo, xo) with yo*stride - pad <= yi < yo*stride - pad + k
    const int yo_hi = min(ho - 1, (yi + pad) / stride), xo_hi = min(wo - 1, (xi + pad) / stride);
    for (int yo = yo_hi; yo >= 0 && yo * stride - pad + k > yi; --yo) {
      for (int xo = xo_hi; xo >= 0 && xo * stride - pad + k > xi; --xo) {
        float mx[8], d[8];
        const size_t o = (((size_t)img * ho + yo) * wo + xo) * c + g * 8;
        unpack8(__ldg(reinterpret_cast<const uint4*>(y + o)), mx);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + o)), d);
        bool cand[8];
        bool any = false;
#pragma unroll
        for (int q = 0; q < 8; ++q) { cand[q] = me[q] == mx[q]; any |= cand[q]; }
        if (!any) continue;
        // earlier elements of the window (scan order) holding the same maximum win the tie
        const int y0 = yo * stride - pad, x0 = xo * stride - pad;
        for (int r = 0; r < k; ++r) {
          const int yy = y0 + r;
          if (yy < 0 || yy >= h || yy > yi) continue;
          for (int s2 = 0; s2 < k; ++s2) {
            const int xx = x0 + s2;
            if (xx < 0 || xx >= w) continue;
            if (yy == yi && xx >= xi) break;
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)img * h + yy) * w + xx) * c + g * 8)), f);
#pragma unroll
            for (int q = 0; q < 8; ++q) if (f[q] == mx[q]) cand[q] = false;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) if (cand[q]) acc[q] += d[q];
      }
    }
    *reinterpret_cast<uint4*>(dx + (((size_t)img * h + yi) * w + xi) * c + g * 8) = pack8(acc);
  }
}

// mean((a-b)^2) over two NHWC bf16 maps (F.mse_loss on VGG features, loss.py:81): out[0] += sum (a-b)^2 * inv_numel;
// da = grad_scale * 2 (a-b) * inv_numel  (bf16; NULL to skip)
__global__ void mse_feat_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n8,
                                float inv_numel, float grad_scale, float* __restrict__ out, __nv_bfloat16* __restrict__ da) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a) + i), x);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b) + i), y);
#pragma unroll
    for (int q = 0; q < 8; ++q) { const float d = x[q] - y[q]; acc = fmaf(d, d, acc); x[q] = 2.f * d * inv_numel * grad_scale; }
    if (da) reinterpret_cast<uint4*>(da)[i] = pack8(x);
  }
  __shared__ float s_red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += s_red[wi];
    atomicAdd(out, t * inv_numel);
  }
}

// LPIPS tap (lpips.py normalize_tensor / spatial_average / lin): per pixel na = fa/(|fa|+eps), nb likewise,
// v = sum_c w_c (na_c - nb_c)^2;  out[img] += v / hw;  dfa = grad_scale/hw * d v / d fa  (bf16; NULL to skip).
// One warp per pixel, lanes stride the 16-byte channel groups.
__global__ void lpips_tap_kernel(const __nv_bfloat16* __restrict__ fa, const __nv_bfloat16* __restrict__ fb, int n, long long hw,
                                 int c, const float* __restrict__ lin_w, float grad_scale, float* __restrict__ out,
                                 __nv_bfloat16* __restrict__ dfa) {
  const int G = c / 8;
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)n * hw;
  const float inv_hw = 1.f / (float)hw;
  for (long long p = warp_id; p < total; p += nwarps) {
    const int img = (int)(p / hw);
    float sa = 0.f, sb = 0.f;
    for (int g = lane; g < G; g += 32) {
      float x[8], y[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(fa + (size_t)p * c + g * 8)), x);
      unpack8(__ldg(reinterpret_cast<const uint4*>(fb + (size_t)p * c + g * 8)), y);
#pragma unroll
      for (int q = 0; q < 8; ++q) { sa = fmaf(x[q], x[q], sa); sb = fmaf(y[q], y[q], sb); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    const float ra = sqrtf(sa), rb = sqrtf(sb);
    const float ia = 1.f / (ra + 1e-10f), ib = 1.f / (rb + 1e-10f);
    float v = 0.f, dot = 0.f;      // dot = sum_c q_c fa_c, q_c = 2 w_c d_c
    for (int g = lane; g < G; g += 32) {
      float x[8], y[8], wv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(fa + (size_t)p * c + g * 8)), x);
      unpack8(__ldg(reinterpret_cast<const uint4*>(fb + (size_t)p * c + g * 8)), y);
      load8f(lin_w + g * 8, wv);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float d = x[q] * ia - y[q] * ib;
        v = fmaf(wv[q] * d, d, v);
        dot = fmaf(2.f * wv[q] * d, x[q], dot);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); dot += __shfl_xor_sync(0xffffffffu, dot, o); }
    if (lane == 0) atomicAdd(out + img, v * inv_hw);
    if (dfa) {
      const float k2 = ra > 0.f ? dot * ia * ia / ra : 0.f;     // (sum q.fa) / (r s^2)
      const float gs = grad_scale * inv_hw;
      for (int g = lane; g < G; g += 32) {
        float x[8], y[8], wv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(fa + (size_t)p * c + g * 8)), x);
        unpack8(__ldg(reinterpret_cast<const uint4*>(fb + (size_t)p * c + g * 8)), y);
        load8f(lin_w + g * 8, wv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float d = x[q] * ia - y[q] * ib;
          x[q] = gs * (2.f * wv[q] * d * ia - x[q] * k2);
        }
        *reinterpret_cast<uint4*>(dfa + (size_t)p * c + g * 8) = pack8(x);
      }
    }
  }
}

// Transpose of adb_stem_pack: dx[i,c,y,x] (+)= scale[c] * sum_{r,s} dcols[i, yo, xo, (r*kw+s)*3+c] over the output
// positions (yo, xo) whose tap (r, s) reads input pixel (y, x).  kh == 1: horizontal taps only (yo = y).
__global__ void stem_unpack_kernel(const __nv_bfloat16* __restrict__ dcols, int n, int h, int w, int ho, int wo, int kh, int kw,
                                   int pad, int stride, int kp, float s0, float s1, float s2, int accumulate,
                                   float* __restrict__ dx) {
  const long long total = (long long)n * h * w;
  const int sh = kh > 1 ? stride : 1, ph = kh > 1 ? pad : 0;
  const size_t plane = (size_t)h * w;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int xi = (int)(t % w);
    const int yi = (int)((t / w) % h);
    const int img = (int)(t / ((long long)w * h));
    float acc[3] = {0.f, 0.f, 0.f};
    for (int r = 0; r < kh; ++r) {
      const int ynum = yi + ph - r;
      if (ynum < 0 || ynum % sh) continue;
      const int yo = ynum / sh;
      if (yo >= ho) continue;
      for (int s = 0; s < kw; ++s) {
        const int xnum = xi + pad - s;
        if (xnum < 0 || xnum % stride) continue;
        const int xo = xnum / stride;
        if (xo >= wo) continue;
        const __nv_bfloat16* q = dcols + (((size_t)img * ho + yo) * wo + xo) * kp + (r * kw + s) * 3;
        acc[0] += __bfloat162float(q[0]); acc[1] += __bfloat162float(q[1]); acc[2] += __bfloat162float(q[2]);
      }
    }
    float* o = dx + (size_t)img * 3 * plane + (size_t)yi * w + xi;
    const float sc[3] = {s0, s1, s2};
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = (accumulate ? o[c * plane] : 0.f) + acc[c] * sc[c];
  }
}

// ------------------------------------------------------------------ classifier head / pooling backward (classifier.py:72-97)
// dx[n,h,w,c] (bf16) = dfeat[n,c] * scale  — backward of the global average pool (scale = 1/hw)
__global__ void broadcast_hw_kernel(const float* __restrict__ dfeat, int n, long long hw, int c, float scale,
                                    __nv_bfloat16* __restrict__ dx) {
  const int G = c / 8;
  const long long total = (long long)n * hw * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long p = t / G;
    const int img = (int)(p / hw);
    float f[8];
    load8f(dfeat + (size_t)img * c + g * 8, f);
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] *= scale;
    *reinterpret_cast<uint4*>(dx + (size_t)p * c + g * 8) = pack8(f);
  }
}
// out = a * b, zeroed where gate <= 0 (dropout masks, ReLU backward of the head)
__global__ void mul_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gate, long long n,
                               float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = a[i] * (b ? b[i] : 1.f);
    out[i] = (gate && gate[i] <= 0.f) ? 0.f : v;
  }
}
// nn.Linear backward on small fp32 matrices: dx[n][fin] = dy W ; dw[fout][fin] = dy^T x ; db[fout] = sum_n dy
__global__ void linear_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, int n, int fin, int fout,
                                     float* __restrict__ dx) {
  const int i = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < fin; k += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int o = 0; o < fout; ++o) acc = fmaf(dy[(size_t)i * fout + o], w[(size_t)o * fin + k], acc);
    dx[(size_t)i * fin + k] = acc;
  }
}
__global__ void linear_bwd_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, int n, int fin, int fout,
                                     float* __restrict__ dw, float* __restrict__ db) {
  const int o = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < fin; k += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(dy[(size_t)i * fout + o], x[(size_t)i * fin + k], acc);
    dw[(size_t)o * fin + k] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += dy[(size_t)i * fout + o];
    db[o] = s;
  }
}

// backward of the 2x2/2 average pool (DenseNet transition): dx[2y+a, 2x+b] = dy[y, x] / 4
__global__ void avgpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int pitch_dy, int n, int h, int w, int c,
                                      __nv_bfloat16* __restrict__ dx, int pitch_dx) {
  const int G = c / 8, ho = h / 2, wo = w / 2;
  const long long total = (long long)n * h * w * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xi = (int)(p % w); p /= w;
    const int yi = (int)(p % h);
    const int img = (int)(p / h);
    float f[8];
    const int yo = yi >> 1, xo = xi >> 1;
    if (yo < ho && xo < wo) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (((size_t)img * ho + yo) * wo + xo) * pitch_dy + g * 8)), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] *= 0.25f;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = 0.f;
    }
    *reinterpret_cast<uint4*>(dx + (((size_t)img * h + yi) * w + xi) * pitch_dx + g * 8) = pack8(f);
  }
}

// Re-pack a parameter for the kernels after an optimizer step: out[i] = bf16(src[idx[i]]) (0 where idx[i] < 0).  The
// index map is derived once per packing (training/autograd.py), so a step re-packs every weight with one launch each
// instead of a dozen torch ops.
__global__ void gather_cast_kernel(const float* __restrict__ src, const int* __restrict__ idx, long long n,
                                   __nv_bfloat16* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = __ldg(idx + i);
    out[i] = __float2bfloat16(k >= 0 ? __ldg(src + k) : 0.f);
  }
}

// The same for MANY packings in one launch (a training step re-packs ~600 tensors; one launch each made the re-pack
// launch-bound: 4.6 ms for 0.3 GB).  jobs[j] = {src, idx, out, n, first_block}; a block finds its job by binary search.
struct GatherJob { const float* src; const int* idx; __nv_bfloat16* out; long long n; long long first_block; };
constexpr int kGatherPerBlock = 256 * 8;
__global__ void gather_cast_multi_kernel(const GatherJob* __restrict__ jobs, int njobs) {
  int lo = 0, hi = njobs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (jobs[mid].first_block <= (long long)blockIdx.x) lo = mid; else hi = mid;
  }
  const GatherJob J = jobs[lo];
  const long long base = ((long long)blockIdx.x - J.first_block) * kGatherPerBlock;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const long long i = base + u * 256 + threadIdx.x;
    if (i < J.n) {
      const int k = __ldg(J.idx + i);
      J.out[i] = __float2bfloat16(k >= 0 ? __ldg(J.src + k) : 0.f);
    }
  }
}

// backward of nn.UpsamplingBilinear2d(scale) (align_corners=True; medium_intensity.py:146,151, high_intensity.py:171,173)
// as a gather: a source pixel collects from every destination pixel whose two-tap footprint touches it, with the weights
// recomputed exactly as the forward computes them.  dy may be a channel slice of a wider map (pitch_dy, c_off).
__global__ void upsample_bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int pitch_dy, int c_off, int n, int h, int w,
                                             int c, int scale, __nv_bfloat16* __restrict__ dx) {
  const int G = c / 8, ho = h * scale, wo = w * scale;
  const float rh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f;
  const float rw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
  const long long total = (long long)n * h * w * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xi = (int)(p % w); p /= w;
    const int yi = (int)(p % h);
    const int img = (int)(p / h);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    // destination rows whose source coordinate lies in (yi - 1, yi + 1)
    const int yo_lo = rh > 0.f ? max(0, (int)floorf((float)(yi - 1) / rh)) : 0;
    const int yo_hi = rh > 0.f ? min(ho - 1, (int)ceilf((float)(yi + 1) / rh)) : ho - 1;
    const int xo_lo = rw > 0.f ? max(0, (int)floorf((float)(xi - 1) / rw)) : 0;
    const int xo_hi = rw > 0.f ? min(wo - 1, (int)ceilf((float)(xi + 1) / rw)) : wo - 1;
    for (int yo = yo_lo; yo <= yo_hi; ++yo) {
      const float sy = rh * (float)yo;
      const int y0 = (int)sy;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
      const float ly = sy - (float)y0;
      const float wy = (y0 == yi ? 1.f - ly : 0.f) + (y1 == yi ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int xo = xo_lo; xo <= xo_hi; ++xo) {
        const float sx = rw * (float)xo;
        const int x0 = (int)sx;
        const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float lx = sx - (float)x0;
        const float wx = (x0 == xi ? 1.f - lx : 0.f) + (x1 == xi ? lx : 0.f);
        if (wx == 0.f) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (((size_t)img * ho + yo) * wo + xo) * pitch_dy + c_off + g * 8)), f);
        const float wgt = wy * wx;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(wgt, f[q], acc[q]);
      }
    }
    *reinterpret_cast<uint4*>(dx + (((size_t)img * h + yi) * w + xi) * c + g * 8) = pack8(acc);
  }
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------ Adam (torch.optim.Adam semantics: L2 weight decay added to the gradient)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float gr = fmaf(wd, w, g[i] * grad_scale);
    const float mm = fmaf(b1, m[i], (1.f - b1) * gr);
    const float vv = fmaf(b2, v[i], (1.f - b2) * gr * gr);
    m[i] = mm; v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = w - (lr / bc1) * (mm / denom);
  }
}

// Per-parameter ("segment") Adam: torch.optim.Adam skips a parameter whose .grad is None — no moment decay, no weight decay,
// no step count — so a branch that received no sample (HardRouter joint training) does not drift.  `live[s]` > 0 when
// any rank produced a gradient for segment s this step (the flags ride at the tail of the all-reduced bucket).
__global__ void adam_seg_prep_kernel(const float* __restrict__ live, int* __restrict__ step, float* __restrict__ bc1,
                                     float* __restrict__ bc2s, int n_seg, float b1, float b2) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg || !(live[s] > 0.f)) return;
  const int st = step[s] + 1;
  step[s] = st;
  bc1[s] = 1.f - powf(b1, (float)st);
  bc2s[s] = sqrtf(1.f - powf(b2, (float)st));
}

__global__ void adam_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n4, float lr, float b1, float b2, float eps, float wd, float grad_scale,
                                const long long* __restrict__ seg_off, int n_seg, const float* __restrict__ live,
                                const float* __restrict__ bc1, const float* __restrict__ bc2s) {
  // segments start on multiples of 4 elements, so one lookup serves a float4
  for (long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
    const long long i = i4 * 4;
    int lo = 0, hi = n_seg;                     // last segment whose offset is <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid;
    }
    if (!(__ldg(live + lo) > 0.f)) continue;
    const float c1 = __ldg(bc1 + lo), c2 = __ldg(bc2s + lo);
    const float4 w4 = *reinterpret_cast<const float4*>(p + i), g4 = *reinterpret_cast<const float4*>(g + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
    float w[4] = {w4.x, w4.y, w4.z, w4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = fmaf(wd, w[k], gg[k] * grad_scale);
      mm[k] = fmaf(b1, mm[k], (1.f - b1) * gr);
      vv[k] = fmaf(b2, vv[k], (1.f - b2) * gr * gr);
      const float denom = sqrtf(vv[k]) / c2 + eps;
      w[k] = w[k] - (lr / c1) * (mm[k] / denom);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(p + i) = make_float4(w[0], w[1], w[2], w[3]);
  }
}

// slot chunks of stat_fold_kernel: enough CTAs to keep the fold short, at least 64 slots per chunk
inline int stat_fold_chunks(int64_t slots) { return (int)std::max<int64_t>(1, std::min<int64_t>(256, slots / 64)); }

}  // namespace

#define ADB_BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define ADB_BFM(p) reinterpret_cast<__nv_bfloat16*>(p)

extern "C" {

int64_t adb_bn_scratch_floats(int64_t pixels, int32_t c) {
  return (int64_t)red_blocks(pixels, c, sm_count()) * 2 * c + 3 * (int64_t)c;
}

int adb_bn_train_stats(const void* z, int64_t pixels, int32_t c, int32_t pitch, const float* gamma, const float* beta, float eps,
                       float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, float* scratch,
                       float* mean, float* rstd, float* scale, float* shift, void* stream) {
  ADB_REQUIRE(z && gamma && beta && scratch && mean && rstd && scale && shift, "adb_bn_train_stats: null pointer");
  ADB_REQUIRE(pixels > 0 && c > 0 && c % 8 == 0 && c <= 2048 && pitch >= c && pitch % 8 == 0, "adb_bn_train_stats: bad shape (pixels=%lld c=%d pitch=%d)", (long long)pixels, c, pitch);
  const int sms = sm_count();
  const int nb = red_blocks(pixels, c, sms);
  const int G = c / 8, PY = std::max(1, kRedThreads / G);
  cudaStream_t st = (cudaStream_t)stream;
  chan_reduce_kernel<0><<<nb, kRedThreads, (size_t)PY * 2 * c * sizeof(float), st>>>(ADB_BF(z), pitch, nullptr, 0, nullptr, 0, pixels, c, 0,
                                                                                   nullptr, nullptr, nullptr, 0, scratch);
  ADB_CUDA_OK(cudaGetLastError());
  bn_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(scratch, nb, c, (double)pixels, gamma, beta, eps, momentum, running_mean, running_var,
                                                      reinterpret_cast<long long*>(num_batches_tracked), mean, rstd, scale, shift);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int64_t adb_bn_stat_scratch_floats(int64_t slots, int32_t c) {
  return (int64_t)stat_fold_chunks(slots) * 2 * c;
}

int adb_bn_finalize_stats(const float* stat, int64_t slots, int32_t cpitch, int64_t pixels, int32_t c, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          int64_t* num_batches_tracked, float* scratch, float* mean, float* rstd, float* scale, float* shift,
                          void* stream) {
  ADB_REQUIRE(stat && gamma && beta && scratch && mean && rstd && scale && shift, "adb_bn_finalize_stats: null pointer");
  ADB_REQUIRE(slots > 0 && pixels > 0 && c > 0 && c <= cpitch && c <= 2048, "adb_bn_finalize_stats: bad shape (slots=%lld pixels=%lld c=%d cpitch=%d)",
              (long long)slots, (long long)pixels, c, cpitch);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = stat_fold_chunks(slots);
  stat_fold_kernel<<<dim3((c + 31) / 32, chunks), dim3(32, 8), 0, st>>>(stat, slots, cpitch, c, scratch);
  ADB_CUDA_OK(cudaGetLastError());
  bn_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(scratch, chunks, c, (double)pixels, gamma, beta, eps, momentum, running_mean, running_var,
                                                      reinterpret_cast<long long*>(num_batches_tracked), mean, rstd, scale, shift);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_affine_act(const void* z, int32_t pitch_z, int64_t pixels, int32_t c, const float* scale, const float* shift,
                   const void* residual, int32_t pitch_r, int32_t act, void* y, int32_t pitch_y, void* stream) {
  ADB_REQUIRE(z && scale && shift && y, "adb_affine_act: null pointer");
  ADB_REQUIRE(pixels > 0 && c > 0 && c % 8 == 0 && pitch_z % 8 == 0 && pitch_y % 8 == 0 && (!residual || pitch_r % 8 == 0), "adb_affine_act: bad shape");
  affine_act_kernel<<<ew_grid(pixels, c / 8, sm_count()), kEwThreads, 0, (cudaStream_t)stream>>>(
      ADB_BF(z), pitch_z, pixels, c, scale, shift, ADB_BF(residual), pitch_r, act, ADB_BFM(y), pitch_y);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_bn_bwd(const void* dy, int32_t pitch_dy, const void* y, int32_t pitch_y, const void* z, int32_t pitch_z, int64_t pixels,
               int32_t c, int32_t act, const float* gamma, const float* mean, const float* rstd, float* scratch, void* g_out,
               int32_t pitch_g, void* dz, int32_t pitch_dz, float* dgamma, float* dbeta, int32_t accumulate, void* stream) {
  ADB_REQUIRE(dy && scratch && g_out, "adb_bn_bwd: null pointer");
  ADB_REQUIRE((gamma != nullptr) == (z != nullptr), "adb_bn_bwd: gamma and z go together (both NULL for a bias-only layer)");
  ADB_REQUIRE(!gamma || (mean && rstd && dz), "adb_bn_bwd: BatchNorm backward needs mean/rstd/dz");
  ADB_REQUIRE(pixels > 0 && c > 0 && c % 8 == 0 && c <= 2048, "adb_bn_bwd: bad shape");
  const int sms = sm_count();
  const int nb = red_blocks(pixels, c, sms);
  const int G = c / 8, PY = std::max(1, kRedThreads / G);
  cudaStream_t st = (cudaStream_t)stream;
  float* coef = scratch + (size_t)nb * 2 * c;
  chan_reduce_kernel<1><<<nb, kRedThreads, (size_t)PY * 2 * c * sizeof(float), st>>>(ADB_BF(z), pitch_z, ADB_BF(dy), pitch_dy, ADB_BF(y), pitch_y,
                                                                                   pixels, c, act, mean, rstd, ADB_BFM(g_out), pitch_g, scratch);
  ADB_CUDA_OK(cudaGetLastError());
  bn_bwd_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(scratch, nb, c, (double)pixels, gamma, mean, rstd, accumulate, dgamma, dbeta, coef);
  ADB_CUDA_OK(cudaGetLastError());
  if (gamma) {
    bn_bwd_apply_kernel<<<ew_grid(pixels, G, sms), kEwThreads, 0, st>>>(ADB_BF(g_out), pitch_g, ADB_BF(z), pitch_z, pixels, c, coef,
                                                                            ADB_BFM(dz), pitch_dz);
    ADB_CUDA_OK(cudaGetLastError());
  }
  return ADB_OK;
}

int adb_bn_relu_bwd(const void* dy, int32_t pitch_dy, const void* z, int32_t pitch_z, int64_t pixels, int32_t c, const float* scale,
                    const float* shift, const float* gamma, const float* mean, const float* rstd, float* scratch, void* dz,
                    int32_t pitch_dz, int32_t dz_accumulate, float* dgamma, float* dbeta, int32_t accumulate, void* stream) {
  ADB_REQUIRE(dy && z && scale && shift && gamma && mean && rstd && scratch && dz, "adb_bn_relu_bwd: null pointer");
  ADB_REQUIRE(pixels > 0 && c > 0 && c % 8 == 0 && c <= 2048 && pitch_dy % 8 == 0 && pitch_z % 8 == 0 && pitch_dz % 8 == 0,
              "adb_bn_relu_bwd: bad shape (pixels=%lld c=%d)", (long long)pixels, c);
  ADB_REQUIRE(!dz_accumulate || dz != dy, "adb_bn_relu_bwd: an accumulated dz cannot alias dy");
  const int sms = sm_count();
  const int nb = red_blocks(pixels, c, sms);
  const int G = c / 8, PY = std::max(1, kRedThreads / G);
  cudaStream_t st = (cudaStream_t)stream;
  float* coef = scratch + (size_t)nb * 2 * c;
  bn_relu_bwd_reduce_kernel<<<nb, kRedThreads, (size_t)PY * 2 * c * sizeof(float), st>>>(ADB_BF(z), pitch_z, ADB_BF(dy), pitch_dy, pixels, c,
                                                                                       scale, shift, mean, rstd, scratch);
  ADB_CUDA_OK(cudaGetLastError());
  bn_bwd_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(scratch, nb, c, (double)pixels, gamma, mean, rstd, accumulate, dgamma, dbeta, coef);
  ADB_CUDA_OK(cudaGetLastError());
  const int grid = ew_grid(pixels, G, sms);
  if (dz_accumulate)
    bn_relu_bwd_apply_kernel<true><<<grid, kEwThreads, 0, st>>>(ADB_BF(dy), pitch_dy, ADB_BF(z), pitch_z, pixels, c, scale, shift, coef, ADB_BFM(dz), pitch_dz);
  else
    bn_relu_bwd_apply_kernel<false><<<grid, kEwThreads, 0, st>>>(ADB_BF(dy), pitch_dy, ADB_BF(z), pitch_z, pixels, c, scale, shift, coef, ADB_BFM(dz), pitch_dz);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_add_bf16(void* a, int32_t pitch_a, const void* b, int32_t pitch_b, int64_t pixels, int32_t c, void* stream) {
  ADB_REQUIRE(a && b && pixels > 0 && c % 8 == 0 && pitch_a % 8 == 0 && pitch_b % 8 == 0, "adb_add_bf16: bad arguments");
  add_bf16_kernel<<<ew_grid(pixels, c / 8, sm_count()), kEwThreads, 0, (cudaStream_t)stream>>>(ADB_BFM(a), pitch_a, ADB_BF(b), pitch_b, pixels, c);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_img_head_fwd(const void* z, int32_t pitch, const float* x, const float* guidance, const float* alpha, int32_t mode,
                     int32_t act, int32_t n, int32_t h, int32_t w, float* out, void* stream) {
  ADB_REQUIRE(z && x && out && pitch >= 8 && pitch % 8 == 0, "adb_img_head_fwd: bad arguments");
  ADB_REQUIRE(mode != ADB_IMG_GUIDED || guidance, "adb_img_head_fwd: GUIDED needs guidance");
  ADB_REQUIRE(mode != ADB_IMG_BLEND || alpha, "adb_img_head_fwd: BLEND needs alpha");
  const long long hw = (long long)h * w;
  img_head_fwd_kernel<<<grid_for((long long)n * hw, 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(ADB_BF(z), pitch, x, guidance, alpha, mode, act, n, hw, out);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_img_head_bwd(const float* dout, const void* z, int32_t pitch, const float* x, const float* guidance, const float* alpha,
                     int32_t mode, int32_t act, int32_t n, int32_t h, int32_t w, void* dz, float* dguidance, float* red4,
                     void* stream) {
  ADB_REQUIRE(dout && z && x && dz && red4 && pitch >= 8 && pitch % 8 == 0, "adb_img_head_bwd: bad arguments");
  ADB_REQUIRE(mode != ADB_IMG_GUIDED || (guidance && dguidance), "adb_img_head_bwd: GUIDED needs guidance/dguidance");
  ADB_REQUIRE(mode != ADB_IMG_BLEND || alpha, "adb_img_head_bwd: BLEND needs alpha");
  const long long hw = (long long)h * w;
  cudaStream_t st = (cudaStream_t)stream;
  ADB_CUDA_OK(cudaMemsetAsync(red4, 0, 4 * sizeof(float), st));
  img_head_bwd_kernel<<<grid_for((long long)n * hw, 256, sm_count(), 8), 256, 0, st>>>(dout, ADB_BF(z), pitch, x, guidance, alpha, mode, act, n, hw,
                                                                                        ADB_BFM(dz), dguidance, red4);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_dot_head_fwd(const void* y, int32_t pitch, int32_t c, const float* w, const float* b, int64_t pixels, float* g, void* stream) {
  ADB_REQUIRE(y && w && b && g && c % 8 == 0 && pitch >= c, "adb_dot_head_fwd: bad arguments");
  dot_head_fwd_kernel<<<grid_for(pixels, 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(ADB_BF(y), pitch, c, w, b, pixels, g);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_dot_head_bwd(const float* dg, const float* g, const void* y, int32_t pitch, int32_t c, const float* w, int64_t pixels, void* dy,
                     int32_t pitch_dy, float* red /*[c+1]*/, void* stream) {
  ADB_REQUIRE(dg && g && y && w && dy && red, "adb_dot_head_bwd: null pointer");
  ADB_REQUIRE(c == 16 || c == 32, "adb_dot_head_bwd: c must be 16 or 32 (got %d)", c);
  cudaStream_t st = (cudaStream_t)stream;
  ADB_CUDA_OK(cudaMemsetAsync(red, 0, (size_t)(c + 1) * sizeof(float), st));
  const int grid = grid_for(pixels, 256, sm_count(), 4);
  if (c == 16) dot_head_bwd_kernel<16><<<grid, 256, 0, st>>>(dg, g, ADB_BF(y), pitch, w, pixels, ADB_BFM(dy), pitch_dy, red);
  else dot_head_bwd_kernel<32><<<grid, 256, 0, st>>>(dg, g, ADB_BF(y), pitch, w, pixels, ADB_BFM(dy), pitch_dy, red);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

static int attn_chunks(long long hw, int c, int n, int sms) {
  const int G = c / 8, PY = std::max(1, kRedThreads / G);
  long long chunks = (hw + PY * 8 - 1) / (PY * 8);
  chunks = std::min<long long>(chunks, std::max(1, sms * 4 / std::max(1, n)));
  return (int)std::max<long long>(1, chunks);
}

int64_t adb_attn_bwd_scratch_floats(int32_t n, int32_t h, int32_t w, int32_t c) {
  const long long hw = (long long)h * w;
  // d_pre[n*hw] | amax[n*hw] (int) | d_stats[n*hw*2] | partials[n*chunks*c] | d_avg[n*c] | d_max[n*c] | pos[n*c] (int)
  return (int64_t)n * hw * 4 + (int64_t)n * attn_chunks(hw, c, n, sm_count()) * c + 3LL * n * c;
}

int adb_attn_bwd(const void* dy, const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const float* pool, const float* gate,
                 const float* stats, const float* spatial, const float* w1, const float* w2, int32_t c_red, const float* w_spatial,
                 float* scratch, void* dx, float* dw1, float* dw2, float* dw_spatial, void* stream) {
  ADB_REQUIRE(dy && x && pool && gate && stats && spatial && w1 && w2 && w_spatial && scratch && dx && dw1 && dw2 && dw_spatial,
              "adb_attn_bwd: null pointer");
  ADB_REQUIRE(n > 0 && h > 0 && w > 0 && c % 8 == 0 && c <= 2048 && c_red > 0, "adb_attn_bwd: bad shape");
  const long long hw = (long long)h * w;
  ADB_REQUIRE(hw < (1LL << 31), "adb_attn_bwd: map too large");
  const int sms = sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = attn_chunks(hw, c, n, sms);
  float* d_pre = scratch;
  int* amax = reinterpret_cast<int*>(scratch + (size_t)n * hw);
  float* d_stats = scratch + 2 * (size_t)n * hw;
  float* partials = scratch + 4 * (size_t)n * hw;
  float* d_avg = partials + (size_t)n * chunks * c;
  float* d_max = d_avg + (size_t)n * c;
  int* pos = reinterpret_cast<int*>(d_max + (size_t)n * c);
  const int G = c / 8;
  const long long total = (long long)n * hw;
#define ADB_BWPIX(LP_, ML_) attn_bwd_pixel_kernel<LP_, ML_><<<grid_for(total * (LP_), 256, sms, 8), 256, 0, st>>>( \
    ADB_BF(dy), ADB_BF(x), n, hw, c, gate, spatial, d_pre, amax)
  if (G <= 4) ADB_BWPIX(4, 1);
  else if (G <= 8) ADB_BWPIX(8, 1);
  else if (G <= 16) ADB_BWPIX(16, 1);
  else if (G <= 32) ADB_BWPIX(32, 1);
  else if (G <= 64) ADB_BWPIX(32, 2);
  else if (G <= 128) ADB_BWPIX(32, 4);
  else ADB_BWPIX(32, 8);
#undef ADB_BWPIX
  ADB_CUDA_OK(cudaGetLastError());
  ADB_CUDA_OK(cudaMemsetAsync(dw_spatial, 0, 98 * sizeof(float), st));
  {
    dim3 block(kBwTW, kBwTH), grid((w + kBwTW - 1) / kBwTW, (h + kBwTH - 1) / kBwTH, n);
    attn_bwd_stencil_kernel<<<grid, block, 0, st>>>(d_pre, stats, h, w, w_spatial, d_stats, dw_spatial);
    ADB_CUDA_OK(cudaGetLastError());
  }
  fill_int_kernel<<<grid_for((long long)n * c, 256, sms, 1), 256, 0, st>>>(pos, (long long)n * c, 0x7fffffff);
  {
    const int PY = std::max(1, kRedThreads / G);
    dim3 grid(chunks, n);
    attn_bwd_main_kernel<<<grid, kRedThreads, (size_t)PY * c * sizeof(float), st>>>(ADB_BF(dy), ADB_BF(x), hw, c, gate, spatial, d_stats, amax, pool,
                                                                                   ADB_BFM(dx), partials, pos);
    ADB_CUDA_OK(cudaGetLastError());
  }
  ADB_CUDA_OK(cudaMemsetAsync(dw1, 0, (size_t)c_red * c * sizeof(float), st));
  ADB_CUDA_OK(cudaMemsetAsync(dw2, 0, (size_t)c_red * c * sizeof(float), st));
  attn_bwd_gate_kernel<<<n, 256, (size_t)(3 * c + 4 * c_red) * sizeof(float), st>>>(partials, chunks, pool, 1.f / (float)hw, c, c_red, gate, w1, w2,
                                                                                    d_avg, d_max, dw1, dw2);
  ADB_CUDA_OK(cudaGetLastError());
  attn_bwd_pool_kernel<<<grid_for(total * G, 256, sms, 16), 256, 0, st>>>(ADB_BFM(dx), n, hw, c, d_avg, d_max, pos);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_blend3_bwd(const float* dout, const float* y0, const float* y1, const float* y2, const float* weights, float temperature,
                   int32_t b, int64_t chw, float* dy0, float* dy1, float* dy2, float* dweights /*[b][3]*/, float* dlogits /*nullable*/,
                   void* stream) {
  ADB_REQUIRE(dout && y0 && y1 && y2 && weights && dy0 && dy1 && dy2 && dweights, "adb_blend3_bwd: null pointer");
  ADB_REQUIRE(b > 0 && chw > 0 && chw % 4 == 0, "adb_blend3_bwd: chw must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  ADB_CUDA_OK(cudaMemsetAsync(dweights, 0, (size_t)b * 3 * sizeof(float), st));
  const long long chw4 = chw / 4;
  dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((chw4 + 255) / 256, (long long)sm_count() * 8 / std::max(1, b) + 1)), (unsigned)b);
  blend3_bwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(y0),
                                          reinterpret_cast<const float4*>(y1), reinterpret_cast<const float4*>(y2), weights, chw4,
                                          reinterpret_cast<float4*>(dy0), reinterpret_cast<float4*>(dy1), reinterpret_cast<float4*>(dy2), dweights);
  ADB_CUDA_OK(cudaGetLastError());
  if (dlogits) {
    ADB_REQUIRE(temperature > 0.f, "adb_blend3_bwd: dlogits needs temperature > 0");
    softmax3_bwd_kernel<<<(b + 127) / 128, 128, 0, st>>>(weights, dweights, 1.f / temperature, b, dlogits);
    ADB_CUDA_OK(cudaGetLastError());
  }
  return ADB_OK;
}

int adb_image_affine(const float* x, int32_t n, int32_t h, int32_t w, const float* scale3_host, const float* shift3_host, float* y,
                     void* stream) {
  ADB_REQUIRE(x && y && scale3_host && shift3_host && n > 0, "adb_image_affine: bad arguments");
  const long long hw = (long long)h * w, total = 3LL * n * hw;
  image_affine_kernel<<<grid_for(total, 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(x, hw, total, scale3_host[0], scale3_host[1],
                                                                                           scale3_host[2], shift3_host[0], shift3_host[1], shift3_host[2], y);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_maxpool_fwd(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride, int32_t pad, void* y, void* stream) {
  ADB_REQUIRE(x && y && n > 0 && c % 8 == 0 && k >= 2 && k <= 4 && stride >= 1 && pad >= 0 && pad < k, "adb_maxpool_fwd: bad arguments");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  ADB_REQUIRE(ho > 0 && wo > 0, "adb_maxpool_fwd: empty output");
  maxpool_fwd_kernel<<<grid_for((long long)n * ho * wo * (c / 8), 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(
      ADB_BF(x), n, h, w, c, k, stride, pad, ho, wo, ADB_BFM(y));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_maxpool_bwd(const void* dy, const void* x, const void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride,
                    int32_t pad, void* dx, void* stream) {
  ADB_REQUIRE(dy && x && y && dx && n > 0 && c % 8 == 0 && k >= 2 && k <= 4 && stride >= 1 && pad >= 0 && pad < k, "adb_maxpool_bwd: bad arguments");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  maxpool_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8), 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(
      ADB_BF(dy), ADB_BF(x), ADB_BF(y), n, h, w, c, k, stride, pad, ho, wo, ADB_BFM(dx));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_mse_feat(const void* a, const void* b, int64_t numel, float grad_scale, float* out /* += */, void* da, void* stream) {
  ADB_REQUIRE(a && b && out && numel > 0 && numel % 8 == 0, "adb_mse_feat: numel must be a positive multiple of 8");
  mse_feat_kernel<<<grid_for(numel / 8, 256, sm_count(), 8), 256, 0, (cudaStream_t)stream>>>(ADB_BF(a), ADB_BF(b), numel / 8, 1.f / (float)numel,
                                                                                          grad_scale, out, ADB_BFM(da));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_lpips_tap(const void* fa, const void* fb, int32_t n, int32_t h, int32_t w, int32_t c, const float* lin_w, float grad_scale,
                  float* out /*[n] += */, void* dfa, void* stream) {
  ADB_REQUIRE(fa && fb && lin_w && out && n > 0 && c % 8 == 0, "adb_lpips_tap: bad arguments");
  const long long hw = (long long)h * w;
  lpips_tap_kernel<<<grid_for((long long)n * hw * 32, 256, sm_count(), 8), 256, 0, (cudaStream_t)stream>>>(ADB_BF(fa), ADB_BF(fb), n, hw, c, lin_w,
                                                                                                        grad_scale, out, ADB_BFM(dfa));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_stem_unpack(const void* dcols, int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t pad, int32_t stride, int32_t kp,
                    const float* scale3_host, int32_t accumulate, float* dx, void* stream) {
  ADB_REQUIRE(dcols && dx && scale3_host && n > 0 && kp % 8 == 0 && kp >= 3 * kh * (kh > 1 ? kw : 1) && kp >= 3 * kw, "adb_stem_unpack: bad arguments");
  const int wo = (w + 2 * pad - kw) / stride + 1;
  const int ho = kh > 1 ? (h + 2 * pad - kh) / stride + 1 : h;
  stem_unpack_kernel<<<grid_for((long long)n * h * w, 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(
      ADB_BF(dcols), n, h, w, ho, wo, kh > 1 ? kh : 1, kw, pad, stride, kp, scale3_host[0], scale3_host[1], scale3_host[2], accumulate, dx);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_broadcast_hw(const float* dfeat, int32_t n, int32_t h, int32_t w, int32_t c, float scale, void* dx, void* stream) {
  ADB_REQUIRE(dfeat && dx && n > 0 && c % 8 == 0, "adb_broadcast_hw: bad arguments");
  const long long hw = (long long)h * w;
  broadcast_hw_kernel<<<grid_for((long long)n * hw * (c / 8), 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(dfeat, n, hw, c, scale, ADB_BFM(dx));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_mul_f32(const float* a, const float* b, const float* gate, int64_t n, float* out, void* stream) {
  ADB_REQUIRE(a && out && n > 0, "adb_mul_f32: bad arguments");
  mul_f32_kernel<<<grid_for(n, 256, sm_count(), 8), 256, 0, (cudaStream_t)stream>>>(a, b, gate, n, out);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_linear_bwd(const float* x, const float* w, const float* dy, int32_t n, int32_t fin, int32_t fout, float* dx /*nullable*/,
                   float* dw, float* db, void* stream) {
  ADB_REQUIRE(x && w && dy && dw && db && n > 0 && fin > 0 && fout > 0, "adb_linear_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    linear_bwd_dx_kernel<<<dim3((fin + 127) / 128, n), 128, 0, st>>>(dy, w, n, fin, fout, dx);
    ADB_CUDA_OK(cudaGetLastError());
  }
  linear_bwd_dw_kernel<<<dim3((fin + 127) / 128, fout), 128, 0, st>>>(dy, x, n, fin, fout, dw, db);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_avgpool2x2_bwd(const void* dy, int32_t pitch_dy, int32_t n, int32_t h, int32_t w, int32_t c, void* dx, int32_t pitch_dx,
                       void* stream) {
  ADB_REQUIRE(dy && dx && n > 0 && c % 8 == 0 && pitch_dy % 8 == 0 && pitch_dx % 8 == 0, "adb_avgpool2x2_bwd: bad arguments");
  avgpool2x2_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8), 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(
      ADB_BF(dy), pitch_dy, n, h, w, c, ADB_BFM(dx), pitch_dx);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_gather_cast(const float* src, const int32_t* idx, int64_t n, void* out, void* stream) {
  ADB_REQUIRE(src && idx && out && n > 0, "adb_gather_cast: bad arguments");
  gather_cast_kernel<<<grid_for(n, 256, sm_count(), 8), 256, 0, (cudaStream_t)stream>>>(src, idx, n, ADB_BFM(out));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_gather_cast_multi(const void* jobs_dev, int32_t njobs, int64_t total_blocks, void* stream) {
  ADB_REQUIRE(jobs_dev && njobs > 0 && total_blocks > 0 && total_blocks < (1LL << 31), "adb_gather_cast_multi: bad arguments");
  gather_cast_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const GatherJob*>(jobs_dev), njobs);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_upsample_bilinear_bwd(const void* dy, int32_t pitch_dy, int32_t c_off, int32_t n, int32_t h, int32_t w, int32_t c, int32_t scale,
                              void* dx, void* stream) {
  ADB_REQUIRE(dy && dx && n > 0 && c % 8 == 0 && pitch_dy % 8 == 0 && c_off % 8 == 0 && scale >= 1 && scale <= 8, "adb_upsample_bilinear_bwd: bad arguments");
  upsample_bilinear_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8), 256, sm_count(), 16), 256, 0, (cudaStream_t)stream>>>(
      ADB_BF(dy), pitch_dy, c_off, n, h, w, c, scale, ADB_BFM(dx));
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t step, float grad_scale, void* stream) {
  ADB_REQUIRE(param && grad && exp_avg && exp_avg_sq && numel > 0 && step >= 1, "adb_adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(numel, 256, sm_count(), 8), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps,
                                                                                     weight_decay, bc1, sqrtf(bc2), grad_scale);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_adam_step_segments(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr, float beta1,
                           float beta2, float eps, float weight_decay, float grad_scale, const int64_t* seg_offsets, int32_t n_seg,
                           const float* seg_live, int32_t* seg_step, float* seg_bc1, float* seg_bc2s, void* stream) {
  ADB_REQUIRE(param && grad && exp_avg && exp_avg_sq && numel > 0 && numel % 4 == 0, "adb_adam_step_segments: bad arguments (numel %% 4 == 0)");
  ADB_REQUIRE(seg_offsets && n_seg > 0 && seg_live && seg_step && seg_bc1 && seg_bc2s, "adb_adam_step_segments: null segment tables");
  adam_seg_prep_kernel<<<(n_seg + 255) / 256, 256, 0, (cudaStream_t)stream>>>(seg_live, seg_step, seg_bc1, seg_bc2s, n_seg, beta1, beta2);
  adam_seg_kernel<<<grid_for(numel / 4, 256, sm_count(), 8), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, numel / 4, lr, beta1, beta2, eps, weight_decay, grad_scale,
      reinterpret_cast<const long long*>(seg_offsets), n_seg, seg_live, seg_bc1, seg_bc2s);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

}  // extern "C"
