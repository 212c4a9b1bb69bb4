// guard_fp32.cu — the route guard: HDEN logits of near-tie images recomputed in fp32 on the CUDA cores.
//
// HDEN runs in bf16 with fp32 accumulation; its logits carry an error of a few 1e-4 (bounded at 2e-2 in the tests).  The
// reference routes on fp32 logits (routing.py:41-43: classifier(x) -> torch.argmax), so an image whose two largest logits
// are closer than that error could take a different branch here than in the reference — and then its whole output is
// different.  The guard removes that case without a host round trip:
//   1. guard_flags_kernel     lists, in ascending order, the batch rows whose bf16 top-2 gap is below eps (device count);
//   2. the kernels below      re-run the classifier trunk for the listed rows only, in fp32 storage and fp32 FMA arithmetic
//                             (NHWC fp32 maps, direct implicit GEMM on the CUDA cores, same op order as the reference
//                             graph: conv -> BatchNorm affine -> ReLU / residual / pools), `cap` rows per pass;
//   3. guard_scatter_kernel   writes the fp32 logits over the bf16 ones before adb_route takes the argmax.
// Every kernel of a pass reads the pass's first list position from a device cursor and the list length from the device
// count, and exits at once when its rows are not live, so the host can enqueue passes blindly (or, better, enqueue ONE
// CUDA graph whose WHILE node repeats the pass while rows remain: adb_guard_graph_*).  With a trained HDEN the list is
// almost always empty and the guard costs one flag kernel + one graph launch per batch.
//
// Reference arithmetic replaced: models/classifier.py:80-97 (torchvision resnet18/34 and densenet121 trunks + head) at
// fp32, for the rows the guard selects; models/routing.py:41-43 consumes the result.
#include "adb_host.h"
#include <algorithm>

namespace {

constexpr int kBM = 64;       // output pixels per CTA tile
constexpr int kBK = 16;       // K (tap, input channel) slice per step
constexpr int kThreads = 256;

struct Live {                 // the rows of this pass: list positions [cursor, cursor + live)
  const int* index;           // flag list (batch rows)
  const int* count;           // list length
  const int* cursor;          // first list position of this pass
  int cap;
};
__device__ __forceinline__ int live_rows(const Live& L) { return max(0, min(L.cap, *L.count - *L.cursor)); }

struct F32Conv {
  Live live;
  const float* x;             // NHWC fp32 [cap][H][W][in_pitch]; or (in_nchw) unused
  const float* const* x_slot; // in_nchw: *x_slot = the NCHW fp32 image batch; row = live.index[cursor + i]
  int in_nchw;
  int H, W, Cin, in_pitch;
  int kh, kw, stride, pad, Ho, Wo;
  const float* w;             // [kh*kw][Cin][Cout]
  int Cout;
  const float* pre_scale; const float* pre_shift;     // nullable: the conv sees relu(x*pre_scale[c] + pre_shift[c]) (zero padding AFTER it)
  const float* post_scale; const float* post_shift;   // nullable: y = acc*post_scale[co] + post_shift[co]
  int post_relu;
  const float* residual; int res_pitch;               // nullable: added before the ReLU (BasicBlock identity)
  float* y; int out_pitch, out_c_off;
};

template <int BN>
__global__ void __launch_bounds__(kThreads)
f32_conv_kernel(const F32Conv P) {
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][BN];
  const int live = live_rows(P.live);
  const long long px_total = (long long)live * P.Ho * P.Wo;
  const long long px0 = (long long)blockIdx.x * kBM;
  if (px0 >= px_total) return;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  constexpr int TN = BN / 16;                 // output channels per thread
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // the pixel this thread loads for the A tile
  const int lp = tid >> 2, lk = (tid & 3) * 4;
  const long long lpx = px0 + lp;
  const bool lvalid = lpx < px_total;
  int li = 0, lho = 0, lwo = 0;
  if (lvalid) {
    li = (int)(lpx / ((long long)P.Ho * P.Wo));
    const int r = (int)(lpx - (long long)li * P.Ho * P.Wo);
    lho = r / P.Wo; lwo = r - lho * P.Wo;
  }
  const float* img_nchw = nullptr;
  if (P.in_nchw && lvalid) img_nchw = *P.x_slot + (size_t)P.live.index[*P.live.cursor + li] * P.Cin * P.H * P.W;
  const int ktot = P.kh * P.kw * P.Cin;
  const bool fast = (P.Cin % kBK) == 0 && !P.in_nchw;     // a K slice lies inside one tap, 16-byte loads
  for (int k0 = 0; k0 < ktot; k0 += kBK) {
    // ---- A tile: [kBK][kBM], pre-activation applied, zero outside the image
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (lvalid) {
      if (fast) {
        const int tap = k0 / P.Cin, c = k0 - tap * P.Cin + lk;
        const int r = tap / P.kw, s = tap - r * P.kw;
        const int hi = lho * P.stride - P.pad + r, wi = lwo * P.stride - P.pad + s;
        if (hi >= 0 && hi < P.H && wi >= 0 && wi < P.W) {
          const float4 v = *reinterpret_cast<const float4*>(P.x + (((size_t)li * P.H + hi) * P.W + wi) * P.in_pitch + c);
          av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
          if (P.pre_scale) {
#pragma unroll
            for (int e = 0; e < 4; ++e) av[e] = fmaxf(fmaf(av[e], P.pre_scale[c + e], P.pre_shift[c + e]), 0.f);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k0 + lk + e;
          if (k < ktot) {
            const int tap = k / P.Cin, c = k - tap * P.Cin;
            const int r = tap / P.kw, s = tap - r * P.kw;
            const int hi = lho * P.stride - P.pad + r, wi = lwo * P.stride - P.pad + s;
            if (hi >= 0 && hi < P.H && wi >= 0 && wi < P.W) {
              float v = P.in_nchw ? img_nchw[((size_t)c * P.H + hi) * P.W + wi]
                                  : P.x[(((size_t)li * P.H + hi) * P.W + wi) * P.in_pitch + c];
              if (P.pre_scale) v = fmaxf(fmaf(v, P.pre_scale[c], P.pre_shift[c]), 0.f);
              av[e] = v;
            }
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) As[lk + e][lp] = av[e];
    // ---- B tile: [kBK][BN]
    for (int i = tid; i < kBK * BN / 4; i += kThreads) {
      const int kk = i / (BN / 4), c4 = (i - kk * (BN / 4)) * 4;
      const int k = k0 + kk, co = n0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < ktot) {
        const float* wp = P.w + (size_t)k * P.Cout + co;
        if (co + 3 < P.Cout && (P.Cout & 3) == 0) v = *reinterpret_cast<const float4*>(wp);
        else {
          if (co < P.Cout) v.x = wp[0];
          if (co + 1 < P.Cout) v.y = wp[1];
          if (co + 2 < P.Cout) v.z = wp[2];
          if (co + 3 < P.Cout) v.w = wp[3];
        }
      }
      *reinterpret_cast<float4*>(&Bs[kk][c4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      float bb[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bb[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long px = px0 + ty * 4 + i;
    if (px >= px_total) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co >= P.Cout) continue;
      float v = acc[i][j];
      if (P.post_scale) v = fmaf(v, P.post_scale[co], P.post_shift[co]);
      if (P.residual) v += P.residual[(size_t)px * P.res_pitch + co];
      if (P.post_relu) v = fmaxf(v, 0.f);
      P.y[(size_t)px * P.out_pitch + P.out_c_off + co] = v;
    }
  }
}

// NHWC fp32 pools over the live rows.  mode 0: 3x3 stride-2 pad-1 max (torchvision stems); mode 1: 2x2 stride-2 average
// (DenseNet transitions).  Output channel pitch / offset let the result land in a dense-block buffer prefix.
__global__ void f32_pool_kernel(Live L, const float* __restrict__ x, int H, int W, int C, int in_pitch, int mode,
                                float* __restrict__ y, int Ho, int Wo, int out_pitch) {
  const int live = live_rows(L);
  const long long total = (long long)live * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float v;
    if (mode == 0) {
      v = -INFINITY;
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          const int hi = ho * 2 - 1 + r, wi = wo * 2 - 1 + s;
          if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = fmaxf(v, x[(((size_t)n * H + hi) * W + wi) * in_pitch + c]);
        }
    } else {
      const float* b = x + (((size_t)n * H + ho * 2) * W + wo * 2) * in_pitch + c;
      v = ((b[0] + b[in_pitch]) + (b[(size_t)W * in_pitch] + b[(size_t)W * in_pitch + in_pitch])) * 0.25f;
    }
    y[(((size_t)n * Ho + ho) * Wo + wo) * out_pitch + c] = v;
  }
}

// feats[i][c] = mean over H*W of act(x) with act = relu(x*scale[c] + shift[c]) when scale is given (DenseNet norm5 + relu
// ahead of adaptive_avg_pool2d), identity otherwise (ResNet).  One CTA per (row, 32 channels), fixed summation order.
__global__ void __launch_bounds__(256)
f32_global_avgpool_kernel(Live L, const float* __restrict__ x, int HW, int C, int pitch, const float* __restrict__ scale,
                          const float* __restrict__ shift, float* __restrict__ feats) {
  __shared__ float part[8][32];
  const int live = live_rows(L);
  const int n = blockIdx.y;
  if (n >= live) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane_px = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C) {
    const float sc = scale ? scale[c] : 1.f, sh = scale ? shift[c] : 0.f;
    for (int p = lane_px; p < HW; p += 8) {
      float v = x[((size_t)n * HW + p) * pitch + c];
      if (scale) v = fmaxf(fmaf(v, sc, sh), 0.f);
      s += v;
    }
  }
  part[lane_px][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    feats[(size_t)n * C + c] = t / (float)HW;
  }
}

// rows whose top-2 logit gap is below eps (or that hold a NaN), ascending, + their count; one CTA (b <= a few thousand)
__global__ void __launch_bounds__(1024)
guard_flags_kernel(const float* __restrict__ logits, int b, int classes, float eps, int* __restrict__ index, int* __restrict__ count,
                   int* __restrict__ cursor) {
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < b; start += 1024) {
    const int i = start + threadIdx.x;
    bool flag = false;
    if (i < b && classes >= 2) {
      float m1 = -INFINITY, m2 = -INFINITY;
      bool nan = false;
      for (int k = 0; k < classes; ++k) {
        const float v = logits[(size_t)i * classes + k];
        nan |= (v != v);
        if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
      }
      flag = nan || (m1 - m2) < eps;
    }
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (flag) index[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += warp_cnt[w]; base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { *count = base; *cursor = 0; }
}

__global__ void guard_scatter_kernel(Live L, const float* __restrict__ logits_c, int classes, float* const* __restrict__ logits_slot) {
  const int live = live_rows(L);
  float* logits = *logits_slot;
  for (int i = threadIdx.x; i < live * classes; i += blockDim.x) {
    const int r = i / classes, k = i - r * classes;
    logits[(size_t)L.index[*L.cursor + r] * classes + k] = logits_c[(size_t)r * classes + k];
  }
}

// end of a pass: move the cursor; inside a graph also tell the WHILE node whether rows remain
__global__ void guard_advance_kernel(const int* count, int* cursor, int cap, cudaGraphConditionalHandle handle, int in_graph) {
  const int c = *cursor + cap;
  *cursor = c;
  if (in_graph) cudaGraphSetConditional(handle, c < *count ? 1u : 0u);
}
__global__ void guard_begin_kernel(const int* count, int* cursor, cudaGraphConditionalHandle handle) {
  *cursor = 0;
  cudaGraphSetConditional(handle, *count > 0 ? 1u : 0u);
}
__global__ void guard_set_slots_kernel(const void** slots, const void* p0, const void* p1) { slots[0] = p0; slots[1] = p1; }

struct GuardGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphConditionalHandle handle = 0;
  cudaStream_t capture_stream = nullptr;
  const int* count = nullptr;
  int* cursor = nullptr;
  int cap = 0;
};

}  // namespace

extern "C" {

int adb_guard_flags(const float* logits, int32_t b, int32_t classes, float eps, int32_t* flag_index, int32_t* flag_count,
                    int32_t* cursor, void* stream) {
  ADB_REQUIRE(logits && flag_index && flag_count && cursor && b > 0 && classes >= 1 && classes <= 64, "adb_guard_flags: bad arguments");
  guard_flags_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, b, classes, eps, flag_index, flag_count, cursor);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_guard_set_slots(const void** slots, const void* images, void* logits, void* stream) {
  ADB_REQUIRE(slots, "adb_guard_set_slots: null slots");
  guard_set_slots_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slots, images, logits);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_f32_conv2d(const adb_f32_conv_desc* d, void* stream) {
  ADB_REQUIRE(d && d->flag_index && d->flag_count && d->cursor && d->cap > 0, "adb_f32_conv2d: the live-row list is required");
  ADB_REQUIRE((d->x != nullptr) != (d->in_nchw != 0) && (!d->in_nchw || d->x_slot), "adb_f32_conv2d: give x (NHWC) or x_slot (NCHW image batch)");
  ADB_REQUIRE(d->w && d->y && d->cin > 0 && d->cout > 0 && d->kh > 0 && d->kw > 0 && d->stride >= 1 && d->pad >= 0, "adb_f32_conv2d: bad geometry");
  ADB_REQUIRE(d->in_nchw || (d->in_pitch >= d->cin && d->in_pitch % 4 == 0), "adb_f32_conv2d: input pitch %d must be a multiple of 4 >= cin", d->in_pitch);
  ADB_REQUIRE((d->pre_scale == nullptr) == (d->pre_shift == nullptr) && (d->post_scale == nullptr) == (d->post_shift == nullptr),
              "adb_f32_conv2d: scale/shift pointers go in pairs");
  F32Conv P;
  P.live = {d->flag_index, d->flag_count, d->cursor, d->cap};
  P.x = d->x; P.x_slot = d->x_slot; P.in_nchw = d->in_nchw;
  P.H = d->h_in; P.W = d->w_in; P.Cin = d->cin; P.in_pitch = d->in_pitch;
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad = d->pad;
  P.Ho = (d->h_in + 2 * d->pad - d->kh) / d->stride + 1;
  P.Wo = (d->w_in + 2 * d->pad - d->kw) / d->stride + 1;
  ADB_REQUIRE(P.Ho > 0 && P.Wo > 0, "adb_f32_conv2d: empty output");
  P.w = d->w; P.Cout = d->cout;
  P.pre_scale = d->pre_scale; P.pre_shift = d->pre_shift; P.post_scale = d->post_scale; P.post_shift = d->post_shift;
  P.post_relu = d->post_relu; P.residual = d->residual; P.res_pitch = d->res_pitch;
  P.y = d->y; P.out_pitch = d->out_pitch; P.out_c_off = d->out_c_off;
  ADB_REQUIRE(d->out_pitch >= d->out_c_off + d->cout, "adb_f32_conv2d: output pitch %d cannot hold channels [%d, %d)", d->out_pitch, d->out_c_off, d->out_c_off + d->cout);
  const long long px_max = (long long)d->cap * P.Ho * P.Wo;
  const bool narrow = d->cout <= 32;
  dim3 grid((unsigned)((px_max + kBM - 1) / kBM), (unsigned)((d->cout + (narrow ? 32 : 64) - 1) / (narrow ? 32 : 64)));
  ADB_REQUIRE(grid.y <= 65535, "adb_f32_conv2d: too many output channels");
  if (narrow) f32_conv_kernel<32><<<grid, kThreads, 0, (cudaStream_t)stream>>>(P);
  else f32_conv_kernel<64><<<grid, kThreads, 0, (cudaStream_t)stream>>>(P);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_f32_pool(const int32_t* flag_index, const int32_t* flag_count, const int32_t* cursor, int32_t cap, const float* x, int32_t h,
                 int32_t w, int32_t c, int32_t in_pitch, int32_t mode, float* y, int32_t out_pitch, void* stream) {
  ADB_REQUIRE(flag_count && cursor && cap > 0 && x && y && h > 0 && w > 0 && c > 0 && in_pitch >= c && out_pitch >= c, "adb_f32_pool: bad arguments");
  ADB_REQUIRE(mode == 0 || (mode == 1 && h % 2 == 0 && w % 2 == 0), "adb_f32_pool: mode 0 = 3x3/2 max, mode 1 = 2x2 average (even H/W)");
  const int ho = mode == 0 ? (h - 1) / 2 + 1 : h / 2, wo = mode == 0 ? (w - 1) / 2 + 1 : w / 2;
  const long long total = (long long)cap * ho * wo * c;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
  f32_pool_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Live{flag_index, flag_count, cursor, cap}, x, h, w, c, in_pitch, mode, y, ho, wo, out_pitch);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_f32_global_avgpool(const int32_t* flag_index, const int32_t* flag_count, const int32_t* cursor, int32_t cap, const float* x,
                           int32_t hw, int32_t c, int32_t pitch, const float* scale, const float* shift, float* feats, void* stream) {
  ADB_REQUIRE(flag_count && cursor && cap > 0 && x && feats && hw > 0 && c > 0 && pitch >= c, "adb_f32_global_avgpool: bad arguments");
  ADB_REQUIRE((scale == nullptr) == (shift == nullptr), "adb_f32_global_avgpool: scale and shift go together");
  f32_global_avgpool_kernel<<<dim3((unsigned)((c + 31) / 32), (unsigned)cap), 256, 0, (cudaStream_t)stream>>>(
      Live{flag_index, flag_count, cursor, cap}, x, hw, c, pitch, scale, shift, feats);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_guard_scatter(const int32_t* flag_index, const int32_t* flag_count, const int32_t* cursor, int32_t cap, const float* logits_c,
                      int32_t classes, float* const* logits_slot, void* stream) {
  ADB_REQUIRE(flag_index && flag_count && cursor && cap > 0 && logits_c && logits_slot && classes > 0, "adb_guard_scatter: bad arguments");
  guard_scatter_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(Live{flag_index, flag_count, cursor, cap}, logits_c, classes, logits_slot);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

int adb_guard_advance(const int32_t* flag_count, int32_t* cursor, int32_t cap, void* stream) {
  ADB_REQUIRE(flag_count && cursor && cap > 0, "adb_guard_advance: bad arguments");
  guard_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_count, cursor, cap, 0, 0);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

// ---- one CUDA graph per (classifier, shape): [cursor = 0; rows remain?] -> WHILE { one fp32 pass; cursor += cap; rows remain? }
int adb_guard_graph_begin(const int32_t* flag_count, int32_t* cursor, int32_t cap, void* stream, void** ctx_out) {
  ADB_REQUIRE(flag_count && cursor && cap > 0 && stream && ctx_out, "adb_guard_graph_begin: bad arguments (a non-default stream is required)");
  GuardGraph* g = new GuardGraph();
  g->count = flag_count; g->cursor = cursor; g->cap = cap; g->capture_stream = (cudaStream_t)stream;
  ADB_CUDA_OK(cudaGraphCreate(&g->graph, 0));
  ADB_CUDA_OK(cudaGraphConditionalHandleCreate(&g->handle, g->graph, 0, 0));
  cudaGraphNode_t begin_node;
  {
    cudaKernelNodeParams kp = {};
    void* args[3] = {(void*)&g->count, (void*)&g->cursor, (void*)&g->handle};
    kp.func = (void*)guard_begin_kernel; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    ADB_CUDA_OK(cudaGraphAddKernelNode(&begin_node, g->graph, nullptr, 0, &kp));
  }
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = g->handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t cond_node;
  ADB_CUDA_OK(cudaGraphAddNode(&cond_node, g->graph, &begin_node, 1, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  ADB_CUDA_OK(cudaStreamBeginCaptureToGraph(g->capture_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  *ctx_out = g;
  return ADB_OK;
}

int adb_guard_graph_end(void* ctx) {
  GuardGraph* g = reinterpret_cast<GuardGraph*>(ctx);
  ADB_REQUIRE(g && g->graph && !g->exec, "adb_guard_graph_end: no capture in progress");
  guard_advance_kernel<<<1, 1, 0, g->capture_stream>>>(g->count, g->cursor, g->cap, g->handle, 1);
  cudaGraph_t body = nullptr;
  ADB_CUDA_OK(cudaStreamEndCapture(g->capture_stream, &body));
  ADB_CUDA_OK(cudaGraphInstantiate(&g->exec, g->graph, 0));
  return ADB_OK;
}

int adb_guard_graph_launch(void* ctx, void* stream) {
  GuardGraph* g = reinterpret_cast<GuardGraph*>(ctx);
  ADB_REQUIRE(g && g->exec, "adb_guard_graph_launch: graph not instantiated");
  ADB_CUDA_OK(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  return ADB_OK;
}

int adb_guard_graph_destroy(void* ctx) {
  GuardGraph* g = reinterpret_cast<GuardGraph*>(ctx);
  if (!g) return ADB_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return ADB_OK;
}

}  // extern "C"
