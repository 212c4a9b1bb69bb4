// pointwise.cu — the HBM-bound kernels of the hot path: stem packing, layout conversion, the AttentionBlock passes,
// pooling, the classifier head, the soft blend and the loss reductions.  All are vectorised (16-byte accesses),
// coalesced, and reduce with warp shuffles; grids are sized in multiples of the SM count.
#include "adb_ptx.cuh"
#include "adb_host.h"
#include <algorithm>
#include <math.h>

namespace {

__device__ __forceinline__ int live_images(int n, const int* n_dev, int n_start) {
  return n_dev ? max(0, min(n, *n_dev - n_start)) : n;
}

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

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

inline int grid_for(long long work_items, int threads, int sm_count, int waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)sm_count * waves;
  return (int)std::max<long long>(1, std::min(blocks, cap));
}

// ------------------------------------------------------------------ stem pack
// out[i,ho,wo,j] (bf16, kp channels), j = (r*kw + s)*3 + c  <-  x[row(i), c, ho*sh + r - ph, wo*stride + s - pad]
// kh == 1: rows are not unrolled (ho = h, ph = 0, sh = 1) — the consumer conv walks the kh row taps itself.
// kh  > 1: full im2col (sh = stride, ph = pad) — the consumer is a 1x1 conv (stride-2 HDEN stems).
__global__ void stem_pack_kernel(const float* __restrict__ x, const int* __restrict__ index, const int* n_dev, int n_start,
                                 int n, int h, int w, int ho, int wo, int kh, int kw, int pad, int stride, int kp,
                                 __nv_bfloat16* __restrict__ out) {
  const int n_eff = live_images(n, n_dev, n_start);
  const int groups = kp / 8;
  const long long total = (long long)n_eff * ho * wo * groups;
  const size_t plane = (size_t)h * w;
  const int sh = kh > 1 ? stride : 1, ph = kh > 1 ? pad : 0;
  const int taps = kh * kw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups);
    long long p = t / groups;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int i = (int)(p / ho);
    const int pos = n_start + i;
    const size_t row = index ? (size_t)index[pos] : (size_t)pos;
    const float* xi = x + row * 3 * plane;
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = g * 8 + q;
      const int tap = j / 3, c = j - 3 * tap;
      const int r = tap / kw, s = tap - r * kw;
      const int yy = yo * sh + r - ph;
      const int xx = xo * stride + s - pad;
      f[q] = (tap < taps && yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(xi + c * plane + (size_t)yy * w + xx) : 0.f;
    }
    *reinterpret_cast<uint4*>(out + (((size_t)i * ho + yo) * wo + xo) * kp + g * 8) = pack8(f);
  }
}

// Tiled variant for the shapes the models use.  All geometry is compile-time, so the tap index math folds to immediate
// shared-memory offsets (the generic kernel above spends ~20 integer instructions per bf16 it writes and is
// instruction-bound at ~0.25 of HBM speed).  Per unit (RO output rows x PX pixels of one image):
//   1. stage the 3-channel input footprint in shared memory, coalesced, eight independent loads in flight per thread;
//   2. thread = (pixel, part): build the pixel's channel groups of this part (compile-time taps) and park them in a
//      padded shared tile (row stride KP*2+16 bytes: conflict-free 16-byte stores);
//   3. copy the tile out in memory order with 16-byte stores (a warp writes 512 contiguous bytes).
template <int KH, int KW, int STRIDE, int KP, int RO, int PARTS>
__global__ void __launch_bounds__(256) stem_pack_tiled_kernel(const float* __restrict__ x, const int* __restrict__ index,
                                                              const int* n_dev, int n_start, int n, int h, int w, int ho, int wo,
                                                              int pad, __nv_bfloat16* __restrict__ out) {
  constexpr int PX = STRIDE == 1 ? 128 : 64;
  constexpr int COLS = (PX - 1) * STRIDE + KW;
  constexpr int G = KP / 8;
  constexpr int GPP = G / PARTS;                  // groups per part
  constexpr int TAPS = KH * KW;
  constexpr int SH = KH > 1 ? STRIDE : 1;
  constexpr int TR = (RO - 1) * SH + KH;          // input rows staged per unit
  constexpr int NE = 3 * TR * COLS;
  constexpr int PIX = RO * PX;
  constexpr int OSTRIDE = KP * 2 + 16;            // bytes per pixel in the padded output tile
  static_assert(PIX * PARTS == 256 && G % PARTS == 0, "one thread per (pixel, part)");
  __shared__ float tile[NE];                      // [3][TR][COLS]
  __shared__ __align__(16) unsigned char otile[PIX * OSTRIDE];
  const int ph = KH > 1 ? pad : 0;
  const int n_eff = live_images(n, n_dev, n_start);
  const int segs = (wo + PX - 1) / PX;
  const int rgroups = (ho + RO - 1) / RO;
  const long long units = (long long)n_eff * rgroups * segs;
  const size_t plane = (size_t)h * w;
  const int pp = threadIdx.x % PIX, part = threadIdx.x / PIX;     // part is warp-uniform (PIX is a multiple of 32)
  const int p = pp % PX, ro = pp / PX;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const int seg = (int)(u % segs);
    long long q = u / segs;
    const int yo0 = (int)(q % rgroups) * RO;
    const int i = (int)(q / rgroups);
    const int pos = n_start + i;
    const size_t row = index ? (size_t)index[pos] : (size_t)pos;
    const float* xi = x + row * 3 * plane;
    const int x0 = seg * PX;
    const int xin0 = x0 * STRIDE - pad;
    const int yin0 = yo0 * SH - ph;
    for (int e0 = threadIdx.x; e0 < NE; e0 += 256 * 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = e0 + k * 256;
        const int col = e % COLS;
        const int r = (e / COLS) % TR;
        const int c = e / (COLS * TR);
        const int yy = yin0 + r, xx = xin0 + col;
        v[k] = (e < NE && yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(xi + c * plane + (size_t)yy * w + xx) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = e0 + k * 256;
        if (e < NE) tile[e] = v[k];
      }
    }
    __syncthreads();
    const float* tp = tile + ro * SH * COLS + p * STRIDE;
#pragma unroll
    for (int pt = 0; pt < PARTS; ++pt) {
      if (part == pt) {
#pragma unroll
        for (int gg = 0; gg < GPP; ++gg) {
          const int g = pt * GPP + gg;            // compile-time after unrolling
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = g * 8 + k;
            const int tap = j / 3, c = j - 3 * tap;
            const int r = tap / KW, s2 = tap - r * KW;
            f[k] = tap < TAPS ? tp[(c * TR + r) * COLS + s2] : 0.f;
          }
          *reinterpret_cast<uint4*>(otile + pp * OSTRIDE + g * 16) = pack8(f);
        }
      }
    }
    __syncthreads();
    const int npx = min(PX, wo - x0);
    const int nro = min(RO, ho - yo0);
    for (int t = threadIdx.x; t < nro * npx * G; t += 256) {
      const int g = t % G;
      const int q2 = t / G;
      const int px = q2 % npx, rr = q2 / npx;
      const uint4 val = *reinterpret_cast<const uint4*>(otile + (rr * PX + px) * OSTRIDE + g * 16);
      *reinterpret_cast<uint4*>(out + ((((size_t)i * ho + yo0 + rr) * wo + x0 + px) * KP) + g * 8) = val;
    }
    // the next iteration's staging writes `tile` (last read before the barrier above) and its packing writes `otile`
    // after another barrier, so no third barrier is needed here
  }
}

// ------------------------------------------------------------------ layout converters
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int n, int c, int h, int w, int pitch,
                                    __nv_bfloat16* __restrict__ out) {
  const int groups = pitch / 8;
  const size_t plane = (size_t)h * w;
  const long long total = (long long)n * plane * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const size_t pix = t % plane;
    long long r = t / plane;
    const int g = (int)(r % groups);
    const int i = (int)(r / groups);
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ch = g * 8 + q;
      f[q] = ch < c ? __ldg(x + ((size_t)i * c + ch) * plane + pix) : 0.f;
    }
    *reinterpret_cast<uint4*>(out + ((size_t)i * plane + pix) * pitch + g * 8) = pack8(f);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int n, int c, int h, int w, int pitch,
                                    float* __restrict__ out) {
  const int groups = (c + 7) / 8;
  const size_t plane = (size_t)h * w;
  const long long total = (long long)n * plane * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const size_t pix = t % plane;
    long long r = t / plane;
    const int g = (int)(r % groups);
    const int i = (int)(r / groups);
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((size_t)i * plane + pix) * pitch + g * 8)), f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ch = g * 8 + q;
      if (ch < c) out[((size_t)i * c + ch) * plane + pix] = f[q];
    }
  }
}

// ------------------------------------------------------------------ AttentionBlock pass 1: per-(image, channel) sum and max
// Scratch layout (floats): [n][2][c] result | [n] block tickets (int) | [n][chunks][2][c] per-block partials.
// Deterministic: blocks publish partials, the last block of an image (ticket) folds them in chunk order.
__global__ void pool_init_kernel(int* tickets, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) tickets[t] = 0;
}

// block = G x PY threads (G = c/8 channel groups); grid = (chunks, n).  Each thread streams its 8 channels down a pixel
// chunk with 16-byte loads (a warp reads whole pixel rows: coalesced), then the PY partials meet in shared memory.
__global__ void attn_pool_kernel(const __nv_bfloat16* __restrict__ x, int n, long long hw, int c, const int* n_dev,
                                 int n_start, int pix_per_block, float* __restrict__ pool, int* __restrict__ tickets,
                                 float* __restrict__ partials) {
  const int n_eff = live_images(n, n_dev, n_start);
  const int img = blockIdx.y;
  if (img >= n_eff) return;
  const int chunks = gridDim.x;
  const int G = c / 8;
  const int PY = blockDim.x / G;
  const int g = threadIdx.x % G, py = threadIdx.x / G;
  extern __shared__ float sm[];  // [PY][c] sums then [PY][c] maxes
  __shared__ int s_last;
  float s[8], m[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { s[q] = 0.f; m[q] = -INFINITY; }
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = min(hw, p0 + pix_per_block);
  if (py < PY) {
    const __nv_bfloat16* base = x + (size_t)img * hw * c + g * 8;
    // four independent 16-byte loads in flight per thread; the accumulation order over pixels is fixed (deterministic)
    for (long long p = p0 + py; p < p1; p += 4LL * PY) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pp = p + (long long)u * PY;
        v[u] = pp < p1 ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)pp * c)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p + (long long)u * PY < p1) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int q = 0; q < 8; ++q) { s[q] += f[q]; m[q] = fmaxf(m[q], f[q]); }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sm[py * c + g * 8 + q] = s[q];
      sm[(PY + py) * c + g * 8 + q] = m[q];
    }
  }
  __syncthreads();
  float* mine = partials + ((size_t)img * chunks + blockIdx.x) * 2 * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float ss = 0.f, mm = -INFINITY;
    for (int r = 0; r < PY; ++r) { ss += sm[r * c + ch]; mm = fmaxf(mm, sm[(PY + r) * c + ch]); }
    mine[ch] = ss;
    mine[c + ch] = mm;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + img, 1) == chunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* all = partials + (size_t)img * chunks * 2 * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float ss = 0.f, mm = -INFINITY;
    for (int k = 0; k < chunks; ++k) { ss += all[(size_t)k * 2 * c + ch]; mm = fmaxf(mm, all[(size_t)k * 2 * c + c + ch]); }
    pool[((size_t)img * 2 + 0) * c + ch] = ss;
    pool[((size_t)img * 2 + 1) * c + ch] = mm;
  }
}

// The same [n][2][c] (sum, max) from the producing conv's epilogue partials (adb_conv_desc.stat_out, stat_mode 2: per image
// `slots` x [2][cpitch], one slot per 32 output pixels) instead of a pass over x: CTA (cx, y, img) folds slot chunk y of the image
// for channels 32cx..32cx+31 in a fixed order and writes out[img][y][2][c].  Run twice (slots -> gridDim.y chunks -> 1).
// block = 32 channels x 8 slot lanes.
__global__ void pool_fold_kernel(const float* __restrict__ stat, int slots, int cpitch, int c, int n, const int* n_dev, int n_start,
                                 float* __restrict__ out) {
  const int img = blockIdx.z;
  if (img >= live_images(n, n_dev, n_start)) return;
  __shared__ float s_part[2][8][32];
  const int ch = blockIdx.x * 32 + threadIdx.x, ly = threadIdx.y;
  const int per = (slots + gridDim.y - 1) / gridDim.y;
  const int k0 = blockIdx.y * per, k1 = min(slots, k0 + per);
  const float* base = stat + (size_t)img * slots * 2 * cpitch;
  float a = 0.f, m = -INFINITY;
  if (ch < c) {
    int k = k0 + ly;
    for (; k + 24 < k1; k += 32) {         // four slots in flight per thread
      float va[4], vm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        va[u] = __ldg(base + (size_t)(k + 8 * u) * 2 * cpitch + ch);
        vm[u] = __ldg(base + (size_t)(k + 8 * u) * 2 * cpitch + cpitch + ch);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a += va[u]; m = fmaxf(m, vm[u]); }
    }
    for (; k < k1; k += 8) {
      a += __ldg(base + (size_t)k * 2 * cpitch + ch);
      m = fmaxf(m, __ldg(base + (size_t)k * 2 * cpitch + cpitch + ch));
    }
  }
  s_part[0][ly][threadIdx.x] = a;
  s_part[1][ly][threadIdx.x] = m;
  __syncthreads();
  if (ly == 0 && ch < c) {
    float t = 0.f, mm = -INFINITY;
    for (int l = 0; l < 8; ++l) { t += s_part[0][l][threadIdx.x]; mm = fmaxf(mm, s_part[1][l][threadIdx.x]); }
    float* o = out + ((size_t)img * gridDim.y + blockIdx.y) * 2 * c;
    o[ch] = t;
    o[c + ch] = mm;
  }
}

// gate[n][c] = sigmoid(W2 relu(W1 avg) + W2 relu(W1 max));  one block per image
__global__ void attn_gate_kernel(const float* __restrict__ pool, int n, float inv_hw, int c, int cr, const int* n_dev,
                                 int n_start, const float* __restrict__ w1, const float* __restrict__ w2,
                                 float* __restrict__ gate) {
  const int n_eff = live_images(n, n_dev, n_start);
  const int img = blockIdx.x;
  if (img >= n_eff) return;
  extern __shared__ float sm[];  // avg[c], mx[c], hid[cr]
  float* avg = sm; float* mx = sm + c; float* hid = sm + 2 * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    avg[ch] = pool[((size_t)img * 2 + 0) * c + ch] * inv_hw;
    mx[ch] = pool[((size_t)img * 2 + 1) * c + ch];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < cr; j += nwarps) {
    float a = 0.f, b = 0.f;
    for (int ch = lane; ch < c; ch += 32) { const float wv = w1[(size_t)j * c + ch]; a = fmaf(wv, avg[ch], a); b = fmaf(wv, mx[ch], b); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) hid[j] = fmaxf(a, 0.f) + fmaxf(b, 0.f);   // W2 is linear: W2 relu(a) + W2 relu(b) = W2 (relu(a)+relu(b))
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float v = 0.f;
    for (int j = 0; j < cr; ++j) v = fmaf(w2[(size_t)ch * cr + j], hid[j], v);
    gate[(size_t)img * c + ch] = 1.f / (1.f + __expf(-v));
  }
}

// pass 2: per pixel, mean and max over channels of x*gate.  LP lanes share a pixel, each lane owns the 16-byte channel
// groups sub, sub+LP, ... (at most ML of them, all loaded before any is used: ML independent loads in flight per lane);
// a segmented shuffle tree finishes the reduction.
template <int LP, int ML>
__global__ void attn_stats_kernel(const __nv_bfloat16* __restrict__ x, int n, long long hw, int c, const int* n_dev,
                                  int n_start, const float* __restrict__ gate, float* __restrict__ stats) {
  const int n_eff = live_images(n, n_dev, n_start);
  const int G = c / 8;
  const int sub = threadIdx.x % LP;
  constexpr int PPW = 32 / LP;   // pixels per warp iteration
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long stride = ((long long)gridDim.x * blockDim.x >> 5) * PPW;
  const long long total = (long long)n_eff * hw;
  const float inv_c = 1.f / (float)c;
  for (long long pw = warp_id * PPW; pw < total; pw += stride) {   // warp-uniform trip count (shuffles below)
    const long long p = pw + (threadIdx.x & 31) / LP;
    const bool live = p < total;
    float s = 0.f, m = -INFINITY;
    if (live) {
      const int img = (int)(p / hw);
      const __nv_bfloat16* px = x + (size_t)p * c;
      const float* gt = gate + (size_t)img * c;
      uint4 v[ML];
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        const int g = sub + k * LP;
        v[k] = g < G ? __ldg(reinterpret_cast<const uint4*>(px + g * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        const int g = sub + k * LP;
        if (g < G) {
          float f[8];
          unpack8(v[k], f);
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gt + g * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(gt + g * 8 + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float t = f[q] * gg[q]; s += t; m = fmaxf(m, t); }
        }
      }
    }
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (live && sub == 0) *reinterpret_cast<float2*>(stats + (size_t)p * 2) = make_float2(s * inv_c, m);
  }
}

// pass 3a: spatial gate map sp[n,h,w] = sigmoid(conv7x7(stats)) — a 2-channel stencil over 8 B/pixel, tiled through
// shared memory (32x16 output pixels + 3-pixel halo per block) so every stats value is read from L2/HBM once.
constexpr int kSpTW = 32, kSpTH = 16;
__global__ void attn_spatial_kernel(const float* __restrict__ stats, int n, int h, int w, const int* n_dev, int n_start,
                                    const float* __restrict__ wsp, float* __restrict__ sp) {
  __shared__ float s_w[98];
  __shared__ float2 s_t[kSpTH + 6][kSpTW + 6];
  const int n_eff = live_images(n, n_dev, n_start);
  const int img = blockIdx.z;
  if (img >= n_eff) return;
  const int tid = threadIdx.y * kSpTW + threadIdx.x;
  for (int i = tid; i < 98; i += kSpTW * kSpTH) s_w[i] = wsp[i];
  const int x0 = blockIdx.x * kSpTW - 3, y0 = blockIdx.y * kSpTH - 3;
  const float2* st = reinterpret_cast<const float2*>(stats) + (size_t)img * h * w;
  for (int i = tid; i < (kSpTH + 6) * (kSpTW + 6); i += kSpTW * kSpTH) {
    const int ty = i / (kSpTW + 6), tx = i - ty * (kSpTW + 6);
    const int yy = y0 + ty, xx = x0 + tx;
    s_t[ty][tx] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(st + (size_t)yy * w + xx) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int px = blockIdx.x * kSpTW + threadIdx.x, py = blockIdx.y * kSpTH + threadIdx.y;
  if (px >= w || py >= h) return;
  float acc = 0.f;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy)
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      const float2 v = s_t[threadIdx.y + dy][threadIdx.x + dx];
      acc = fmaf(s_w[dy * 7 + dx], v.x, acc);          // channel 0: mean over channels
      acc = fmaf(s_w[49 + dy * 7 + dx], v.y, acc);     // channel 1: max over channels
    }
  sp[((size_t)img * h + py) * w + px] = 1.f / (1.f + __expf(-acc));
}

// pass 3b: y = x * gate[c] * sp[pixel] — pure streaming.  grid = (blocks, images); a thread walks 16-byte channel groups of
// its image four at a time (four independent loads in flight), all index arithmetic in 32 bits with the division by the
// group count as one multiply-high (gmul = floor(2^32 / G) + 1, exact for hw * G < 2^32, checked on the host).
__global__ void __launch_bounds__(256)
attn_scale_kernel(const __nv_bfloat16* __restrict__ x, int n, uint32_t hw, int c, uint32_t gmul, const int* n_dev,
                  int n_start, const float* __restrict__ gate, const float* __restrict__ sp, __nv_bfloat16* __restrict__ y) {
  const int img = blockIdx.y;
  if (img >= live_images(n, n_dev, n_start)) return;
  const uint32_t G = (uint32_t)c >> 3;
  const uint32_t total = hw * G;
  const uint4* xi = reinterpret_cast<const uint4*>(x + (size_t)img * hw * c);
  uint4* yi = reinterpret_cast<uint4*>(y + (size_t)img * hw * c);
  const float* spi = sp + (size_t)img * hw;
  const float* gi = gate + (size_t)img * c;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += 4 * stride) {
    uint4 v[4];
    float s[4];
    uint32_t g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t t = t0 + u * stride;
      if (t < total) {
        const uint32_t p = G == 1 ? t : __umulhi(t, gmul);
        g[u] = t - p * G;
        v[u] = __ldg(xi + t);
        s[u] = __ldg(spi + p);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t t = t0 + u * stride;
      if (t < total) {
        float f[8];
        unpack8(v[u], f);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gi + g[u] * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gi + g[u] * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = f[q] * gg[q] * s[u];
        yi[t] = pack8(f);
      }
    }
  }
}

// ------------------------------------------------------------------ pooling for the HDEN backbone
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, int ho, int wo,
                                    __nv_bfloat16* __restrict__ y, int pitch_out) {
  const int G = c / 8;
  const long long total = (long long)n * ho * wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int i = (int)(p / ho);
    float m[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) m[q] = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = yo * 2 - 1 + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = xo * 2 - 1 + dx;
        if (xx < 0 || xx >= w) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)i * h + yy) * w + xx) * c + g * 8)), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], f[q]);
      }
    }
    *reinterpret_cast<uint4*>(y + (((size_t)i * ho + yo) * wo + xo) * pitch_out + g * 8) = pack8(m);
  }
}

// k x k / stride k max pool (nn.MaxPool2d(k, k): medium_intensity.py:144,149; high_intensity.py:164,167), NHWC bf16.
// One thread per (output pixel, 8-channel group): k*k independent 16-byte loads, coalesced along channels.
template <int K>
__global__ void maxpool_kxk_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, const int* n_dev,
                                   int n_start, __nv_bfloat16* __restrict__ y) {
  const int n_eff = live_images(n, n_dev, n_start);
  const int G = c / 8, ho = h / K, wo = w / K;
  const long long total = (long long)n_eff * ho * wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int i = (int)(p / ho);
    uint4 v[K * K];
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
      for (int q = 0; q < K; ++q)
        v[r * K + q] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)i * h + yo * K + r) * w + xo * K + q) * c + g * 8));
    float m[8];
    unpack8(v[0], m);
#pragma unroll
    for (int k = 1; k < K * K; ++k) {
      float f[8];
      unpack8(v[k], f);
#pragma unroll
      for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], f[q]);
    }
    *reinterpret_cast<uint4*>(y + (((size_t)i * ho + yo) * wo + xo) * c + g * 8) = pack8(m);
  }
}

// nn.UpsamplingBilinear2d(scale_factor=S) == interpolate(mode='bilinear', align_corners=True) (medium_intensity.py:146,151;
// high_intensity.py:171,173), NHWC bf16 -> channels [c_off, c_off+c) of a [n, h*S, w*S, pitch] buffer (so the multi-scale
// concat of COrunInspiredModel is never materialised).  Source index arithmetic as in ATen: src = dst * (in-1)/(out-1) in
// fp32, i0 = (int)src, lambda = src - i0, i1 = i0 + (i0 < in-1).
__global__ void upsample_bilinear_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, int scale,
                                         const int* n_dev, int n_start, __nv_bfloat16* __restrict__ y, int pitch, int c_off) {
  const int n_eff = live_images(n, n_dev, n_start);
  const int G = c / 8, ho = h * scale, wo = w * scale;
  const float rh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f;
  const float rw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
  const long long total = (long long)n_eff * ho * wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int i = (int)(p / ho);
    const float sy = rh * (float)yo, sx = rw * (float)xo;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const __nv_bfloat16* base = x + (size_t)i * h * w * c + g * 8;
    const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * w + x0) * c));
    const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * w + x1) * c));
    const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y1 * w + x0) * c));
    const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y1 * w + x1) * c));
    float a[8], b[8], cc[8], d[8], o[8];
    unpack8(q00, a); unpack8(q01, b); unpack8(q10, cc); unpack8(q11, d);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = hy * (hx * a[q] + lx * b[q]) + ly * (hx * cc[q] + lx * d[q]);
    *reinterpret_cast<uint4*>(y + (((size_t)i * ho + yo) * wo + xo) * pitch + c_off + g * 8) = pack8(o);
  }
}

// y[p, 0:c] = relu(x[p, 0:c] * scale + shift) — DenseNet pre-activation (norm -> relu ahead of a conv), NHWC bf16
__global__ void affine_relu_kernel(const __nv_bfloat16* __restrict__ x, long long pixels, int c, int pitch_in,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   __nv_bfloat16* __restrict__ y, int pitch_out) {
  const int G = c / 8;
  const long long total = pixels * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long p = t / G;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + (size_t)p * pitch_in + g * 8)), f);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + g * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + g * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + g * 8 + 4));
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = fmaxf(fmaf(f[q], sc[q], sh[q]), 0.f);
    *reinterpret_cast<uint4*>(y + (size_t)p * pitch_out + g * 8) = pack8(f);
  }
}

// 2x2 average pool, stride 2 (DenseNet transition), NHWC bf16
__global__ void avgpool2x2_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, int pitch_in,
                                  __nv_bfloat16* __restrict__ y, int pitch_out) {
  const int G = c / 8, ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long p = t / G;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int i = (int)(p / ho);
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)i * h + 2 * yo + dy) * w + 2 * xo + dx) * pitch_in + g * 8)), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] += f[q];
      }
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] *= 0.25f;
    *reinterpret_cast<uint4*>(y + (((size_t)i * ho + yo) * wo + xo) * pitch_out + g * 8) = pack8(a);
  }
}

__global__ void avgpool_finish_kernel(const float* __restrict__ pool, int n, int c, float inv_hw, float* __restrict__ y) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * c) y[t] = pool[((size_t)(t / c) * 2) * c + (t % c)] * inv_hw;
}

// logits = W2 relu(W1 f + b1) + b2, fp32, one block per image
__global__ void head_mlp_kernel(const float* __restrict__ feat, int f, const float* __restrict__ w1,
                                const float* __restrict__ b1, int hidden, const float* __restrict__ w2,
                                const float* __restrict__ b2, int classes, float* __restrict__ logits) {
  extern __shared__ float sm[];  // feat[f], hid[hidden]
  float* sf = sm; float* hid = sm + f;
  const int img = blockIdx.x;
  for (int i = threadIdx.x; i < f; i += blockDim.x) sf[i] = feat[(size_t)img * f + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < hidden; j += nwarps) {
    float a = 0.f;
    for (int i = lane; i < f; i += 32) a = fmaf(w1[(size_t)j * f + i], sf[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) hid[j] = fmaxf(a + b1[j], 0.f);
  }
  __syncthreads();
  for (int k = warp; k < classes; k += nwarps) {
    float a = 0.f;
    for (int j = lane; j < hidden; j += 32) a = fmaf(w2[(size_t)k * hidden + j], hid[j], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) logits[(size_t)img * classes + k] = a + b2[k];
  }
}

// y = act(W x + b), fp32, one block per row (gate MLP of GatedRouter, routing.py:155-163)
__global__ void linear_kernel(const float* __restrict__ x, int fin, const float* __restrict__ w, const float* __restrict__ b,
                              int fout, int relu, float* __restrict__ y) {
  extern __shared__ float sm[];
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < fin; i += blockDim.x) sm[i] = x[(size_t)row * fin + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < fout; j += nwarps) {
    float a = 0.f;
    for (int i = lane; i < fin; i += 32) a = fmaf(w[(size_t)j * fin + i], sm[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) { a += b ? b[j] : 0.f; y[(size_t)row * fout + j] = relu ? fmaxf(a, 0.f) : a; }
  }
}

// ------------------------------------------------------------------ soft blend
__global__ void blend_weights_kernel(const float* __restrict__ lw, float temperature, int b, float* __restrict__ wout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float a = lw[i * 3 + 0], c = lw[i * 3 + 1], d = lw[i * 3 + 2];
  if (temperature > 0.f) {
    a /= temperature; c /= temperature; d /= temperature;
    const float m = fmaxf(a, fmaxf(c, d));
    a = expf(a - m); c = expf(c - m); d = expf(d - m);
    const float s = a + c + d;
    a /= s; c /= s; d /= s;
  }
  wout[i * 3 + 0] = a; wout[i * 3 + 1] = c; wout[i * 3 + 2] = d;
}

__global__ void blend3_kernel(const float4* __restrict__ y0, const float4* __restrict__ y1, const float4* __restrict__ y2,
                              const float* __restrict__ wts, int b, long long chw4, float4* __restrict__ out) {
  const long long total = (long long)b * chw4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / chw4);
    const float w0 = __ldg(wts + i * 3), w1 = __ldg(wts + i * 3 + 1), w2 = __ldg(wts + i * 3 + 2);
    const float4 a = __ldg(y0 + t), c = __ldg(y1 + t), d = __ldg(y2 + t);
    float4 o;
    // same association order as the reference's three in-place adds into zeros (routing.py:121-127)
    o.x = ((0.f + w0 * a.x) + w1 * c.x) + w2 * d.x;
    o.y = ((0.f + w0 * a.y) + w1 * c.y) + w2 * d.y;
    o.z = ((0.f + w0 * a.z) + w1 * c.z) + w2 * d.z;
    o.w = ((0.f + w0 * a.w) + w1 * c.w) + w2 * d.w;
    out[t] = o;
  }
}

// ------------------------------------------------------------------ losses
__global__ void l1_mse_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, long long numel, float inv,
                                  float* __restrict__ out2) {
  float a = 0.f, s = 0.f;
  const long long n4 = numel / 4;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 u = __ldg(p4 + i), v = __ldg(t4 + i);
    const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
    a += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < numel; i += blockDim.x) { const float d = p[i] - t[i]; a += fabsf(d); s += d * d; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
  __shared__ float sa[32], ss[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sa[warp] = a; ss[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    a = lane < nw ? sa[lane] : 0.f; s = lane < nw ? ss[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
    if (lane == 0) { atomicAdd(out2, a * inv); atomicAdd(out2 + 1, s * inv); }
  }
}

template <int MODE>  // 0: L1, 1: MSE
__global__ void pix_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, long long numel, float k,
                               float* __restrict__ g) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    const float d = p[i] - t[i];
    g[i] = MODE == 0 ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 2.f * d * k;
  }
}

// nn.CrossEntropyLoss semantics (loss.py:177): mean over the rows whose label is not ignore_index (-100); a label outside
// [0, classes) that is not ignore_index raises the library's device error flag (torch raises on the host) and the row is skipped.
__global__ void ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int b, int classes,
                          float grad_scale, float* __restrict__ loss, float* __restrict__ grad, int* __restrict__ err_flag) {
  __shared__ int s_cnt[32];
  int cnt = 0;
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    const long long y = labels[i];
    if (y >= 0 && y < classes) ++cnt;
    else if (y != -100 && err_flag) atomicExch(err_flag, 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  int valid = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) valid += s_cnt[i];
  const float inv = valid > 0 ? 1.f / (float)valid : 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    const float* l = logits + (size_t)i * classes;
    const long long yl = labels[i];
    const bool ok = yl >= 0 && yl < classes;
    const int y = ok ? (int)yl : 0;
    float m = -INFINITY;
    for (int k = 0; k < classes; ++k) m = fmaxf(m, l[k]);
    float s = 0.f;
    for (int k = 0; k < classes; ++k) s += expf(l[k] - m);
    const float lse = m + logf(s);
    if (ok) acc += lse - l[y];
    if (grad)
      for (int k = 0; k < classes; ++k)
        grad[(size_t)i * classes + k] = ok ? (expf(l[k] - lse) - (k == y ? 1.f : 0.f)) * grad_scale * inv : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float sa[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sa[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += sa[i];
    *loss = tot * inv;
  }
}

int sm_count() {
  adbh::DeviceInfo di;
  if (adbh::device_info(&di) != ADB_OK) return 0;
  return di.sm_count;
}

}  // namespace

#define ADB_LAUNCH_OK() ADB_CUDA_OK(cudaGetLastError())

extern "C" {

int adb_stem_pack(const float* x, const int32_t* index, const int32_t* n_dev, int32_t n_start, int32_t n, int32_t h,
                  int32_t w, int32_t kh, int32_t kw, int32_t pad, int32_t stride, int32_t kp, void* out, void* stream) {
  ADB_REQUIRE(x && out && n > 0 && h > 0 && w > 0 && kh >= 1 && kw >= 1, "adb_stem_pack: bad arguments");
  ADB_REQUIRE(kp % 8 == 0 && kp >= kh * kw * 3 && stride >= 1 && stride <= 4, "adb_stem_pack: kp %d must be a multiple of 8 >= 3*kh*kw, stride 1..4", kp);
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const int wo = (w + 2 * pad - kw) / stride + 1;
  const int ho = kh > 1 ? (h + 2 * pad - kh) / stride + 1 : h;
  const long long total = (long long)n * ho * wo * (kp / 8);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  cudaStream_t st = (cudaStream_t)stream;
#define ADB_STEM_TILED(KH_, KW_, S_, KP_, RO_, PARTS_)                                                                   \
  do {                                                                                                                   \
    const int px = (S_) == 1 ? 128 : 64;                                                                                 \
    const long long units = (long long)n * ((ho + (RO_) - 1) / (RO_)) * ((wo + px - 1) / px);                            \
    stem_pack_tiled_kernel<KH_, KW_, S_, KP_, RO_, PARTS_><<<(int)std::min<long long>(units, (long long)sms * 16), 256, 0, st>>>( \
        x, index, n_dev, n_start, n, h, w, ho, wo, pad, o);                                                              \
    ADB_LAUNCH_OK();                                                                                                     \
    return ADB_OK;                                                                                                       \
  } while (0)
  if (kh == 1 && stride == 1 && kw == 3 && kp == 16) ADB_STEM_TILED(1, 3, 1, 16, 2, 1);
  if (kh == 1 && stride == 1 && kw == 7 && kp == 32) ADB_STEM_TILED(1, 7, 1, 32, 2, 1);
  if (kh == 7 && stride == 2 && kw == 7 && kp == 160) ADB_STEM_TILED(7, 7, 2, 160, 1, 4);
#undef ADB_STEM_TILED
  stem_pack_kernel<<<grid_for(total, 256, sms, 16), 256, 0, st>>>(x, index, n_dev, n_start, n, h, w, ho, wo, kh, kw, pad, stride, kp, o);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_nchw_to_nhwc_bf16(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pitch, void* out, void* stream) {
  ADB_REQUIRE(x && out && n > 0 && c > 0 && c_pitch >= c && c_pitch % 8 == 0, "adb_nchw_to_nhwc_bf16: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const long long total = (long long)n * h * w * (c_pitch / 8);
  nchw_to_nhwc_kernel<<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(x, n, c, h, w, c_pitch, reinterpret_cast<__nv_bfloat16*>(out));
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_nhwc_bf16_to_nchw(const void* x, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pitch, float* out, void* stream) {
  ADB_REQUIRE(x && out && n > 0 && c > 0 && c_pitch >= c && c_pitch % 8 == 0, "adb_nhwc_bf16_to_nchw: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const long long total = (long long)n * h * w * ((c + 7) / 8);
  nhwc_to_nchw_kernel<<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, c, h, w, c_pitch, out);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

static void pool_geometry(int n, int h, int w, int sms, long long* chunks, int* ppb) {
  const long long hw = (long long)h * w;
  // enough blocks to fill the device even when a single image of the launch is live (a routed bucket's live count is
  // only known on the device), at least 256 pixels each; n is deliberately not used
  (void)n;
  long long ch = std::max<long long>(1, std::min<long long>((hw + 255) / 256, 2LL * sms));
  *ppb = (int)((hw + ch - 1) / ch);
  *chunks = (hw + *ppb - 1) / *ppb;
}

static int launch_pool(const void* x, int n, int h, int w, int c, const int* n_dev, int n_start, float* pool_buf, cudaStream_t st) {
  ADB_REQUIRE(x && pool_buf && n > 0 && c % 8 == 0 && c / 8 <= 256, "attention/avg pool: channels %d must be a multiple of 8 (<= 2048)", c);
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  long long chunks; int ppb;
  pool_geometry(n, h, w, sms, &chunks, &ppb);
  int* tickets = reinterpret_cast<int*>(pool_buf + (size_t)n * 2 * c);
  float* partials = pool_buf + (size_t)n * 2 * c + ((n + 3) / 4) * 4;
  pool_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(tickets, n);
  const int G = c / 8;
  const int PY = std::max(1, 256 / G);
  const int threads = G * PY;
  dim3 grid((unsigned)chunks, (unsigned)n);
  attn_pool_kernel<<<grid, threads, 2 * PY * c * sizeof(float), st>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, (long long)h * w, c,
                                                                       n_dev, n_start, ppb, pool_buf, tickets, partials);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int64_t adb_pool_scratch_floats(int32_t n, int32_t h, int32_t w, int32_t c) {
  const int sms = sm_count();
  long long chunks; int ppb;
  pool_geometry(n, h, w, sms > 0 ? sms : 148, &chunks, &ppb);
  return (int64_t)n * 2 * c + ((n + 3) / 4) * 4 + (int64_t)n * chunks * 2 * c;
}

int adb_attn_pool(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const int32_t* n_dev, int32_t n_start,
                  float* pool_buf, void* stream) {
  return launch_pool(x, n, h, w, c, n_dev, n_start, pool_buf, (cudaStream_t)stream);
}

// first-level chunk count of pool_fold_kernel
static int pool_fold_chunks(int slots) { return std::max(1, std::min(64, slots / 64)); }

int64_t adb_attn_pool_stat_scratch_floats(int32_t n, int32_t slots_per_image, int32_t c) {
  return (int64_t)n * pool_fold_chunks(slots_per_image) * 2 * c;
}

int adb_attn_pool_from_stats(const float* stat, int32_t n, int32_t slots_per_image, int32_t cpitch, int32_t c, const int32_t* n_dev,
                             int32_t n_start, float* scratch, float* pool_buf, void* stream) {
  ADB_REQUIRE(stat && scratch && pool_buf && n > 0 && slots_per_image > 0 && c > 0 && c <= cpitch, "adb_attn_pool_from_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = pool_fold_chunks(slots_per_image);
  const dim3 block(32, 8);
  pool_fold_kernel<<<dim3((c + 31) / 32, S, n), block, 0, st>>>(stat, slots_per_image, cpitch, c, n, n_dev, n_start, S > 1 ? scratch : pool_buf);
  ADB_LAUNCH_OK();
  if (S > 1) {
    pool_fold_kernel<<<dim3((c + 31) / 32, 1, n), block, 0, st>>>(scratch, S, c, c, n, n_dev, n_start, pool_buf);
    ADB_LAUNCH_OK();
  }
  return ADB_OK;
}

int adb_attn_gate_stats(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const int32_t* n_dev, int32_t n_start,
                        const float* pool_buf, const float* w1, const float* w2, int32_t c_red, float* gate, float* stats,
                        void* stream) {
  ADB_REQUIRE(x && pool_buf && w1 && w2 && gate && stats && n > 0 && c % 8 == 0 && c_red > 0, "adb_attn_gate_stats: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)h * w;
  attn_gate_kernel<<<n, 256, (2 * c + c_red) * sizeof(float), st>>>(pool_buf, n, 1.f / (float)hw, c, c_red, n_dev, n_start, w1, w2, gate);
  ADB_LAUNCH_OK();
  const int G = c / 8;
  const long long total = (long long)n * hw;
  ADB_REQUIRE(G <= 256, "adb_attn_gate_stats: channels %d > 2048", c);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
#define ADB_STATS(LP_, ML_) attn_stats_kernel<LP_, ML_><<<grid_for(total * (LP_), 256, sms, 8), 256, 0, st>>>(xb, n, hw, c, n_dev, n_start, gate, stats)
  if (G <= 8) ADB_STATS(4, 2);
  else if (G <= 12) ADB_STATS(4, 3);
  else if (G <= 16) ADB_STATS(4, 4);
  else if (G <= 24) ADB_STATS(4, 6);
  else if (G <= 48) ADB_STATS(8, 6);
  else if (G <= 128) ADB_STATS(32, 4);
  else ADB_STATS(32, 8);
#undef ADB_STATS
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_attn_apply(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const int32_t* n_dev, int32_t n_start,
                   const float* gate, const float* stats, const float* w_spatial, float* spatial, void* y, void* stream) {
  ADB_REQUIRE(x && gate && stats && w_spatial && spatial && y && n > 0 && c % 8 == 0, "adb_attn_apply: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((w + kSpTW - 1) / kSpTW, (h + kSpTH - 1) / kSpTH, n), block(kSpTW, kSpTH);
  attn_spatial_kernel<<<grid, block, 0, st>>>(stats, n, h, w, n_dev, n_start, w_spatial, spatial);
  const long long per_img = (long long)h * w * (c / 8);
  const uint32_t G = (uint32_t)c / 8;
  ADB_REQUIRE(per_img * G < (1LL << 32), "adb_attn_apply: %lld channel groups per image exceed the 32-bit index range", per_img);
  const uint32_t gmul = G <= 1 ? 0u : (uint32_t)((1ULL << 32) / G) + 1u;
  // ~16 waves over the whole batch, each thread taking four groups per trip
  const int bx = (int)std::max<long long>(1, std::min<long long>((per_img + 1023) / 1024, ((long long)sms * 16 * 8 + n - 1) / n));
  attn_scale_kernel<<<dim3(bx, n), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, (uint32_t)((long long)h * w), c, gmul,
                                                 n_dev, n_start, gate, spatial, reinterpret_cast<__nv_bfloat16*>(y));
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_maxpool3x3s2(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, void* y, int32_t pitch_out, void* stream) {
  ADB_REQUIRE(x && y && n > 0 && c % 8 == 0 && pitch_out >= c && pitch_out % 8 == 0, "adb_maxpool3x3s2: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c, ho, wo, reinterpret_cast<__nv_bfloat16*>(y), pitch_out);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_affine_relu(const void* x, int64_t pixels, int32_t c, int32_t pitch_in, const float* scale, const float* shift,
                    void* y, int32_t pitch_out, void* stream) {
  ADB_REQUIRE(x && y && scale && shift && pixels > 0 && c % 8 == 0 && pitch_in % 8 == 0 && pitch_out % 8 == 0 && pitch_in >= c && pitch_out >= c,
              "adb_affine_relu: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  affine_relu_kernel<<<grid_for(pixels * (c / 8), 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), pixels, c, pitch_in, scale, shift, reinterpret_cast<__nv_bfloat16*>(y), pitch_out);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_maxpool_kxk(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, const int32_t* n_dev, int32_t n_start,
                    void* y, void* stream) {
  ADB_REQUIRE(x && y && n > 0 && c % 8 == 0 && (k == 2 || k == 4) && h % k == 0 && w % k == 0,
              "adb_maxpool_kxk: k must be 2 or 4 and divide H and W, channels a multiple of 8 (k=%d h=%d w=%d c=%d)", k, h, w, c);
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const long long total = (long long)n * (h / k) * (w / k) * (c / 8);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y);
  if (k == 2) maxpool_kxk_kernel<2><<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(xb, n, h, w, c, n_dev, n_start, yb);
  else maxpool_kxk_kernel<4><<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(xb, n, h, w, c, n_dev, n_start, yb);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_upsample_bilinear(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t scale, const int32_t* n_dev,
                          int32_t n_start, void* y, int32_t pitch_out, int32_t c_off, void* stream) {
  ADB_REQUIRE(x && y && n > 0 && c % 8 == 0 && scale >= 1 && pitch_out % 8 == 0 && c_off % 8 == 0 && c_off + c <= pitch_out,
              "adb_upsample_bilinear: bad arguments (c=%d scale=%d pitch=%d c_off=%d)", c, scale, pitch_out, c_off);
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const long long total = (long long)n * h * scale * w * scale * (c / 8);
  upsample_bilinear_kernel<<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c, scale, n_dev, n_start, reinterpret_cast<__nv_bfloat16*>(y), pitch_out, c_off);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_avgpool2x2(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pitch_in, void* y, int32_t pitch_out, void* stream) {
  ADB_REQUIRE(x && y && n > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0 && pitch_in >= c && pitch_out >= c && pitch_in % 8 == 0 && pitch_out % 8 == 0,
              "adb_avgpool2x2: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  avgpool2x2_kernel<<<grid_for(total, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c, pitch_in, reinterpret_cast<__nv_bfloat16*>(y), pitch_out);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

// y[n][c] = mean over h*w; scratch is fp32 [n][2][c].
int adb_global_avgpool(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, float* scratch, float* y, void* stream) {
  ADB_REQUIRE(x && y && scratch && n > 0, "adb_global_avgpool: bad arguments");
  int st = launch_pool(x, n, h, w, c, nullptr, 0, scratch, (cudaStream_t)stream);
  if (st != ADB_OK) return st;
  avgpool_finish_kernel<<<(n * c + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, n, c, 1.f / ((float)h * (float)w), y);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_head_mlp(const float* feat, int32_t n, int32_t f, const float* w1, const float* b1, int32_t hidden,
                 const float* w2, const float* b2, int32_t classes, float* logits, void* stream) {
  ADB_REQUIRE(feat && w1 && b1 && w2 && b2 && logits && n > 0 && f > 0 && hidden > 0 && classes > 0, "adb_head_mlp: bad arguments");
  ADB_REQUIRE((size_t)(f + hidden) * sizeof(float) <= 48 * 1024, "adb_head_mlp: feature dim too large");
  head_mlp_kernel<<<n, 256, (f + hidden) * sizeof(float), (cudaStream_t)stream>>>(feat, f, w1, b1, hidden, w2, b2, classes, logits);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_linear(const float* x, int32_t n, int32_t fin, const float* w, const float* b, int32_t fout, int32_t relu,
               float* y, void* stream) {
  ADB_REQUIRE(x && w && y && n > 0 && fin > 0 && fout > 0 && (size_t)fin * sizeof(float) <= 48 * 1024, "adb_linear: bad arguments");
  linear_kernel<<<n, 256, fin * sizeof(float), (cudaStream_t)stream>>>(x, fin, w, b, fout, relu, y);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_blend3(const float* y0, const float* y1, const float* y2, const float* logits_or_weights, float temperature,
               int32_t b, int64_t chw, float* weights_out, float* out, void* stream) {
  ADB_REQUIRE(y0 && y1 && y2 && logits_or_weights && weights_out && out && b > 0 && chw > 0 && chw % 4 == 0, "adb_blend3: bad arguments (chw must be a multiple of 4)");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  blend_weights_kernel<<<(b + 127) / 128, 128, 0, st>>>(logits_or_weights, temperature, b, weights_out);
  const long long total = (long long)b * (chw / 4);
  blend3_kernel<<<grid_for(total, 256, sms, 16), 256, 0, st>>>(reinterpret_cast<const float4*>(y0), reinterpret_cast<const float4*>(y1),
                                                             reinterpret_cast<const float4*>(y2), weights_out, b, chw / 4, reinterpret_cast<float4*>(out));
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_l1_mse_fwd(const float* pred, const float* target, int64_t numel, float* out2, void* stream) {
  ADB_REQUIRE(pred && target && out2 && numel > 0, "adb_l1_mse_fwd: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  ADB_CUDA_OK(cudaMemsetAsync(out2, 0, 2 * sizeof(float), st));
  l1_mse_fwd_kernel<<<grid_for(numel / 4 + 1, 256, sms, 4), 256, 0, st>>>(pred, target, numel, 1.f / (float)numel, out2);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_l1_bwd(const float* pred, const float* target, int64_t numel, float grad_scale, float* grad, void* stream) {
  ADB_REQUIRE(pred && target && grad && numel > 0, "adb_l1_bwd: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  pix_bwd_kernel<0><<<grid_for(numel, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(pred, target, numel, grad_scale / (float)numel, grad);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_mse_bwd(const float* pred, const float* target, int64_t numel, float grad_scale, float* grad, void* stream) {
  ADB_REQUIRE(pred && target && grad && numel > 0, "adb_mse_bwd: bad arguments");
  const int sms = sm_count();
  if (!sms) return ADB_ERR_NO_DEVICE;
  pix_bwd_kernel<1><<<grid_for(numel, 256, sms, 16), 256, 0, (cudaStream_t)stream>>>(pred, target, numel, grad_scale / (float)numel, grad);
  ADB_LAUNCH_OK();
  return ADB_OK;
}

int adb_ce_fwd_bwd(const float* logits, const int64_t* labels, int32_t b, int32_t classes, float grad_scale, float* loss,
                   float* grad_logits, void* stream) {
  ADB_REQUIRE(logits && labels && loss && b > 0 && classes > 0, "adb_ce_fwd_bwd: bad arguments");
  ce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, reinterpret_cast<const long long*>(labels), b, classes, grad_scale, loss, grad_logits,
                                                 adbh::kernel_err_flag());
  ADB_LAUNCH_OK();
  return ADB_OK;
}

}  // extern "C"
