// conv_wgrad.cu — weight gradient of the implicit-GEMM convolutions for sm_100a (training step).
//
// GEMM view:  dW[m, (tap, c)] = sum_pixels S[pixel, m] * L[pixel + tap, c]
//   S ("small" map): the gradient w.r.t. the conv output (nn.Conv2d), or the layer INPUT of a ConvTranspose2d(4,2,1);
//   L ("large" map): the conv input (one or two concatenated sources), or the output gradient of a ConvTranspose2d.
// The reduction (K) dimension is the PIXEL index, so both operands sit in shared memory exactly as TMA delivers an NHWC
// box — 128-byte rows of 64 channels, one row per pixel — and are consumed as MN-major UMMA operands (instruction
// descriptor a_major = b_major = 1).  One K step of 16 = 16 consecutive pixels of a tile row.
//   M = 128 S-channels (two 64-channel boxes), N = 64 * gpu L-channels, one TMEM accumulator per filter tap of the unit.
// Work unit = (M tile, source, tap set, channel chunk); a tap set is the taps of one filter row (same row/column phase
// of the space-to-depth view for stride 2), which all read ONE halo box {64, TW+ew, 1, TH, 1} through pixel-shifted
// descriptors.  Each unit's pixel range is cut `splits` ways over the grid; a CTA accumulates its range in TMEM and
// writes one fp32 partial [m][k] slab; adb_wgrad's second kernel folds the slabs (deterministic order) into the
// parameter's own layout ([O][I][kh][kw] fp32, the .grad tensor of nn.Conv2d / nn.ConvTranspose2d).
//
// Reference arithmetic replaced: the weight-gradient half of loss.backward() for every nn.Conv2d / nn.ConvTranspose2d
// of models/dehazing/*.py (training/train_dehazing.py:90-92, training/train_joint.py:148-150).
#include "adb_ptx.cuh"
#include "adb_host.h"
#include <algorithm>
#include <string.h>

namespace {

using namespace adb;

constexpr int kWgThreads = 192;     // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5: epilogue
constexpr int kWgMaxSets = 16;
constexpr int kWgMaxSetTaps = 5;
constexpr int kWgMaxStages = 6;

struct WgSet {          // taps of one filter row / phase: one halo box, pixel-shifted views
  int16_t c_mul;        // space-to-depth column phase (channel coordinate += c_mul * pitch)
  int8_t p;             // row phase coordinate
  int8_t dh, dw0;       // halo box origin relative to the tile origin
  uint8_t ntaps;
  uint8_t ddw[kWgMaxSetTaps];     // column shift of each tap inside the halo box
  uint8_t kidx[kWgMaxSetTaps];    // r*kw + s of each tap
};

struct WgK {
  int n, grid_h, grid_w;          // K space: n images of grid_h x grid_w S pixels
  int TW, TH, tiles_w, tiles_h, tw_shift;
  int total_tiles, splits;
  int m_tiles;
  int nsets;
  WgSet sets[kWgMaxSets];
  int chunks0, chunks1;           // channel chunks (of gpu*64 channels) per source
  int c0, c1, pitch0, pitch1;
  int gpu;                        // 64-channel groups per unit (N = 64*gpu)
  int halo_w;
  int s_group_bytes, l_group_bytes, stage_bytes, stages;
  int tmem_cols;
  uint32_t idesc;
  int cs;                         // S channels (rows of dW)
  int ktot;                       // taps * (c0 + c1)
  int m_pad;                      // m_tiles * 128
  float* ws;                      // [splits][m_pad][ktot]
  int* err_flag;
};

// MN-major SWIZZLE_128B operand descriptor: 64 channels (128 B) contiguous per pixel row, 8-pixel groups SBO apart,
// 64-channel groups LBO apart (cute UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// One pixel tile's MMAs as a straight-line run: 8 K steps of 16 pixels x NT taps (one accumulator per tap).
template <int NT>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_base, int N, uint64_t a_hi, uint32_t a_lo, uint64_t b_hi, uint32_t b_lo,
                                            const uint32_t (&b_off)[8], const uint32_t (&t_off)[kWgMaxSetTaps], uint32_t idesc,
                                            uint32_t accumulate) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint64_t a = a_hi | (uint64_t)(a_lo + (uint32_t)kk * 128u);          // + 16 pixel rows of 128 B
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint64_t b = b_hi | (uint64_t)(b_lo + b_off[kk] + t_off[j]);
      umma_bf16(tmem_base + (uint32_t)(j * N), a, b, idesc, accumulate | (uint32_t)kk);
    }
  }
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmL0,
                  const __grid_constant__ CUtensorMap tmL1, const __grid_constant__ WgK P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + (uint32_t)P.stages * (uint32_t)P.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgMaxStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgMaxStages);
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(base_ptr + (size_t)P.stages * P.stage_bytes + 8u * (2 * kWgMaxStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work item: blockIdx.x = unit * splits + split ; unit = ((m_tile * nsets + set) * chunks + chunk)
  const int split = (int)(blockIdx.x % (unsigned)P.splits);
  int unit = (int)(blockIdx.x / (unsigned)P.splits);
  const int nchunks = P.chunks0 + P.chunks1;
  const int chunk = unit % nchunks; unit /= nchunks;
  const int set_i = unit % P.nsets;
  const int m_tile = unit / P.nsets;
  const WgSet S = P.sets[set_i];
  const bool src1 = chunk >= P.chunks0;
  const int ch_base = (src1 ? chunk - P.chunks0 : chunk) * P.gpu * 64;   // first L channel of this unit inside its source
  const int c_src = src1 ? P.c1 : P.c0;
  const int pitch = src1 ? P.pitch1 : P.pitch0;
  const CUtensorMap* tmL = src1 ? &tmL1 : &tmL0;
  const int N = P.gpu * 64;
  // contiguous tile range of this split
  const int per = (P.total_tiles + P.splits - 1) / P.splits;
  const int t_begin = min(P.total_tiles, split * per), t_end = min(P.total_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(tmL);
    for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================================================= producer: S boxes (2 x 64 channels) + L halo boxes (gpu x 64)
    int slot = 0; uint32_t phase = 0;
    const uint32_t tx = 2u * (uint32_t)(128 * 128) + (uint32_t)P.gpu * (uint32_t)(P.halo_w * P.TH * 128);
    for (int t = t_begin; t < t_end; ++t) {
      int q = t;
      const int tw = q % P.tiles_w; q /= P.tiles_w;
      const int th = q % P.tiles_h;
      const int img = q / P.tiles_h;
      const int w0 = tw * P.TW, h0 = th * P.TH;
      mbar_wait(empty_bar(slot), phase ^ 1u, P.err_flag, 11);
      if (elect_one()) {
        const uint32_t sbase = base + (uint32_t)slot * (uint32_t)P.stage_bytes;
        mbar_expect_tx(full_bar(slot), tx);
        tma_load_5d(sbase, &tmS, full_bar(slot), m_tile * 128, w0, 0, h0, img);
        tma_load_5d(sbase + (uint32_t)P.s_group_bytes, &tmS, full_bar(slot), m_tile * 128 + 64, w0, 0, h0, img);
        const uint32_t lbase = sbase + 2u * (uint32_t)P.s_group_bytes;
        for (int g = 0; g < P.gpu; ++g)
          tma_load_5d(lbase + (uint32_t)g * (uint32_t)P.l_group_bytes, tmL, full_bar(slot), S.c_mul * pitch + ch_base + g * 64,
                      w0 + S.dw0, S.p, h0 + S.dh, img);
      }
      __syncwarp();
      if (++slot == P.stages) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    int slot = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    // descriptor halves that never change, and the per-K-step start offsets (16-byte units) inside a stage
    const uint64_t a_hi = make_mnmajor_desc(0, (uint32_t)P.s_group_bytes, 1024u);
    const uint64_t b_hi = make_mnmajor_desc(0, (uint32_t)P.l_group_bytes, 1024u);
    uint32_t b_off[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int px = kk * 16;
      b_off[kk] = (uint32_t)(((px >> P.tw_shift) * P.halo_w + (px & (P.TW - 1))) * 8);
    }
    uint32_t t_off[kWgMaxSetTaps];
#pragma unroll
    for (int j = 0; j < kWgMaxSetTaps; ++j) t_off[j] = (uint32_t)S.ddw[j] * 8u;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full_bar(slot), phase, P.err_flag, 12);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sbase = base + (uint32_t)slot * (uint32_t)P.stage_bytes;
        const uint32_t a_lo = (sbase & 0x3FFFFu) >> 4;
        const uint32_t b_lo = ((sbase + 2u * (uint32_t)P.s_group_bytes) & 0x3FFFFu) >> 4;
        switch (S.ntaps) {
          case 1: issue_wgrad<1>(tmem_base, N, a_hi, a_lo, b_hi, b_lo, b_off, t_off, P.idesc, accumulate); break;
          case 2: issue_wgrad<2>(tmem_base, N, a_hi, a_lo, b_hi, b_lo, b_off, t_off, P.idesc, accumulate); break;
          case 3: issue_wgrad<3>(tmem_base, N, a_hi, a_lo, b_hi, b_lo, b_off, t_off, P.idesc, accumulate); break;
          case 4: issue_wgrad<4>(tmem_base, N, a_hi, a_lo, b_hi, b_lo, b_off, t_off, P.idesc, accumulate); break;
          default: issue_wgrad<5>(tmem_base, N, a_hi, a_lo, b_hi, b_lo, b_off, t_off, P.idesc, accumulate); break;
        }
        umma_commit(empty_bar(slot));
        if (t == t_end - 1) umma_commit(done_bar);
      }
      __syncwarp();
      accumulate = 1;
      if (++slot == P.stages) { slot = 0; phase ^= 1u; }
    }
  } else {
    // ======================================================= epilogue: TMEM -> fp32 partial slab
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int m = m_tile * 128 + q * 32 + lane;
    float* row = P.ws + ((size_t)split * P.m_pad + m) * P.ktot;
    const int ctot = P.c0 + P.c1;
    const int k_src = src1 ? P.c0 : 0;
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0, P.err_flag, 13);
      tc_fence_after();
    }
    for (int j = 0; j < S.ntaps; ++j) {
      float* dst = row + (size_t)S.kidx[j] * ctot + k_src + ch_base;
      for (int c16 = 0; c16 < N / 16; ++c16) {
        if (ch_base + c16 * 16 >= c_src) break;     // (warp-uniform) columns past the source's channels
        float v[16];
        if (t_end > t_begin) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * N + c16 * 16), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (m < P.cs) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(dst + c16 * 16 + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// Fold the per-split slabs and scatter into the parameter layout.
//  layout 0 (OIHW): dw[(m*ctot + c)*taps + tap]                     <- ws[.][m][tap*ctot + c]
//  layout 1 (STEM): the L operand is an adb_stem_pack operand (kp channels = kw_img taps x 3 colours, kh row taps):
//                   dw[((m*3 + c)*kh + r)*kw_img + s]               <- ws[.][m][r*kp + s*3 + c]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int m_pad, int cs, int ktot, int ctot, int taps,
                                    int layout, int kw_img, int accumulate, float* __restrict__ dw) {
  const long long total = (long long)cs * ktot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / ktot), k = (int)(i - (long long)m * ktot);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[((size_t)s * m_pad + m) * ktot + k];
    long long o;
    if (layout == 0) {
      const int tap = k / ctot, c = k - tap * ctot;
      o = ((long long)m * ctot + c) * taps + tap;
    } else {
      const int r = k / ctot, j = k - r * ctot;       // ctot == kp
      const int s = j / 3, c = j - 3 * s;
      if (s >= kw_img) continue;
      o = (((long long)m * 3 + c) * taps + r) * kw_img + s;
    }
    dw[o] = accumulate ? dw[o] + acc : acc;
  }
}

inline int wg_round_up(int a, int b) { return (a + b - 1) / b * b; }

struct WgPlan {
  WgK P;
  int box_w, box_h;
  int units;
  int smem;
  size_t ws_bytes;
};

int wg_plan(const adb_wgrad_desc* d, WgPlan& pl, int sm_count, int max_smem) {
  WgK& P = pl.P;
  memset(&P, 0, sizeof(P));
  ADB_REQUIRE(d != nullptr, "adb_wgrad: null descriptor");
  ADB_REQUIRE(d->grad && d->cg > 0 && d->cg_pitch >= d->cg && d->cg_pitch % 8 == 0 && d->cg % 16 == 0,
              "adb_wgrad: bad small map (cg=%d pitch=%d)", d->cg, d->cg_pitch);
  ADB_REQUIRE(d->act0 && d->c0 > 0 && d->c0 % 16 == 0 && d->c0_pitch >= d->c0 && d->c0_pitch % 8 == 0,
              "adb_wgrad: bad act0 (c0=%d pitch=%d)", d->c0, d->c0_pitch);
  ADB_REQUIRE((d->act1 == nullptr) == (d->c1 == 0), "adb_wgrad: act1/c1 mismatch");
  if (d->act1) ADB_REQUIRE(d->c1 % 16 == 0 && d->c1_pitch >= d->c1 && d->c1_pitch % 8 == 0, "adb_wgrad: bad act1 (c1=%d pitch=%d)", d->c1, d->c1_pitch);
  ADB_REQUIRE(d->n > 0 && d->h_in > 0 && d->w_in > 0, "adb_wgrad: bad n/h/w");
  ADB_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= 64, "adb_wgrad: %dx%d taps unsupported", d->kh, d->kw);

  struct Raw { int c_mul, p, dw, dh, kidx; };
  Raw raw[64];
  const int ntaps = d->kh * d->kw;
  if (d->kind == ADB_CONV_S1) {
    ADB_REQUIRE(d->kh % 2 == 1 && d->kw % 2 == 1, "adb_wgrad: stride-1 convs need odd kernel extents");
    const int pad_h = (d->kh - 1) / 2, pad_w = (d->kw - 1) / 2;
    ADB_REQUIRE(d->pad == std::max(pad_h, pad_w), "adb_wgrad: stride-1 convs must be 'same'");
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) raw[r * d->kw + s] = {0, 0, s - pad_w, r - pad_h, r * d->kw + s};
    P.grid_h = d->h_in; P.grid_w = d->w_in;
  } else if (d->kind == ADB_CONV_S2) {
    ADB_REQUIRE(d->h_in % 2 == 0 && d->w_in % 2 == 0, "adb_wgrad: stride-2 convs need even H/W");
    ADB_REQUIRE((d->h_in + 2 * d->pad - d->kh) / 2 + 1 == d->h_in / 2 && (d->w_in + 2 * d->pad - d->kw) / 2 + 1 == d->w_in / 2,
                "adb_wgrad: stride-2 conv must halve H/W");
    auto fl2 = [](int u) { return (u >= 0) ? u / 2 : -((-u + 1) / 2); };
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) {
        const int u = r - d->pad, v = s - d->pad;
        raw[r * d->kw + s] = {v - 2 * fl2(v), u - 2 * fl2(u), fl2(v), fl2(u), r * d->kw + s};
      }
    P.grid_h = d->h_in / 2; P.grid_w = d->w_in / 2;
  } else {
    return adbh::fail(ADB_ERR_INVALID, "adb_wgrad: kind %d unsupported (a ConvTranspose2d gradient is the stride-2 form with the maps swapped)", d->kind);
  }

  // ---- tap sets: same (c_mul, p, dh)
  bool used[64] = {false};
  int ew = 0;
  P.nsets = 0;
  for (int t0 = 0; t0 < ntaps; ++t0) {
    if (used[t0]) continue;
    int members[64], nm = 0, dw_min = raw[t0].dw, dw_max = raw[t0].dw;
    for (int t1 = t0; t1 < ntaps; ++t1) {
      if (used[t1] || raw[t1].c_mul != raw[t0].c_mul || raw[t1].p != raw[t0].p || raw[t1].dh != raw[t0].dh) continue;
      members[nm++] = t1;
      dw_min = std::min(dw_min, raw[t1].dw); dw_max = std::max(dw_max, raw[t1].dw);
    }
    for (int b = 0; b < nm; b += kWgMaxSetTaps) {     // wide filter rows are cut into sets of <= kWgMaxSetTaps taps
      ADB_REQUIRE(P.nsets < kWgMaxSets, "adb_wgrad: too many tap sets");
      WgSet& s = P.sets[P.nsets++];
      const int cnt = std::min(kWgMaxSetTaps, nm - b);
      int lo = raw[members[b]].dw, hi = lo;
      for (int k = 0; k < cnt; ++k) { lo = std::min(lo, raw[members[b + k]].dw); hi = std::max(hi, raw[members[b + k]].dw); }
      s.c_mul = (int16_t)raw[t0].c_mul; s.p = (int8_t)raw[t0].p; s.dh = (int8_t)raw[t0].dh; s.dw0 = (int8_t)lo;
      s.ntaps = (uint8_t)cnt;
      for (int k = 0; k < cnt; ++k) {
        s.ddw[k] = (uint8_t)(raw[members[b + k]].dw - lo);
        s.kidx[k] = (uint8_t)raw[members[b + k]].kidx;
        used[members[b + k]] = true;
      }
      ew = std::max(ew, hi - lo);
    }
  }
  int max_set_taps = 1;
  for (int i = 0; i < P.nsets; ++i) max_set_taps = std::max(max_set_taps, (int)P.sets[i].ntaps);

  // ---- K tiles: 128 S pixels = TH rows x TW columns, TW >= 16 so a K step of 16 pixels stays inside one row
  int TW = 128;
  while (TW > P.grid_w && TW > 16) TW >>= 1;
  // (maps narrower than 16 pixels still use TW = 16: a K step is then one zero-padded tile row)
  P.TW = TW; P.TH = 128 / TW;
  P.tw_shift = 0;
  while ((1 << P.tw_shift) < TW) ++P.tw_shift;
  P.tiles_w = (P.grid_w + TW - 1) / TW;
  P.tiles_h = (P.grid_h + P.TH - 1) / P.TH;
  const long long tt = (long long)d->n * P.tiles_w * P.tiles_h;
  ADB_REQUIRE(tt < (1LL << 30), "adb_wgrad: too many pixel tiles");
  P.total_tiles = (int)tt;
  P.n = d->n;
  P.halo_w = TW + ew;

  // ---- N: 64-channel groups per unit
  P.c0 = d->c0; P.c1 = d->c1; P.pitch0 = d->c0_pitch; P.pitch1 = d->act1 ? d->c1_pitch : d->c0_pitch;
  const int g0 = (d->c0 + 63) / 64, g1 = (d->c1 + 63) / 64;
  int gpu = std::min(4, 512 / (max_set_taps * 64));
  gpu = std::max(1, std::min(gpu, std::max(g0, g1)));
  P.gpu = gpu;
  P.chunks0 = (g0 + gpu - 1) / gpu;
  P.chunks1 = (g1 + gpu - 1) / gpu;
  const int N = 64 * gpu;
  int cols = 32;
  while (cols < max_set_taps * N) cols <<= 1;
  ADB_REQUIRE(cols <= 512, "adb_wgrad: accumulators do not fit TMEM");
  P.tmem_cols = cols;
  P.idesc = make_idesc_bf16(128u, (uint32_t)N) | (1u << 15) | (1u << 16);   // A and B MN-major

  P.cs = d->cg;
  P.m_tiles = (d->cg + 127) / 128;
  P.m_pad = P.m_tiles * 128;
  P.ktot = ntaps * (d->c0 + d->c1);

  // ---- smem
  P.s_group_bytes = 128 * 128;
  P.l_group_bytes = wg_round_up(P.halo_w * P.TH * 128, 1024);
  P.stage_bytes = 2 * P.s_group_bytes + gpu * P.l_group_bytes;
  const int bar_bytes = 8 * (2 * kWgMaxStages + 1) + 16;
  int stages = (max_smem - 1024 - bar_bytes) / P.stage_bytes;
  stages = std::min(stages, kWgMaxStages);
  ADB_REQUIRE(stages >= 2, "adb_wgrad: pipeline does not fit shared memory");
  P.stages = stages;
  pl.smem = stages * P.stage_bytes + bar_bytes + 1024;
  pl.box_w = P.halo_w; pl.box_h = P.TH;

  // ---- split the pixel range so the grid fills the device for ~2 waves, bounded by the workspace the caller gave
  pl.units = P.m_tiles * P.nsets * (P.chunks0 + P.chunks1);
  int splits = std::max(1, (2 * sm_count + pl.units - 1) / pl.units);
  splits = std::min(splits, std::max(1, P.total_tiles / 4));
  splits = std::min(splits, P.total_tiles);
  const size_t slab = (size_t)P.m_pad * P.ktot * sizeof(float);
  if (d->workspace_bytes > 0) {
    const long long fit = (long long)(d->workspace_bytes / (long long)slab);
    splits = (int)std::max<long long>(0, std::min<long long>(splits, fit));
  }
  P.splits = splits;
  pl.ws_bytes = (size_t)std::max(splits, 1) * slab;
  return ADB_OK;
}


// ====================================================================================================================
// Tap-packed variant for gradients with few channels (cs <= 64: the full-resolution layers of Light / Medium, every
// 3-channel head, the guidance branch).  With M = dZ channels such layers leave half or more of every MMA's rows empty
// and re-read dZ and X once per filter row; here the roles are swapped:
//     D[(tap, ci), co] = sum_pixels X[pixel + tap, ci] * dZ[pixel, co]
//   A (M side) = X: an M tile is TWO 64-channel groups = two filter taps of one channel chunk, addressed through the
//                descriptor's leading-byte-offset (the second tap's view starts (off_b - off_a) pixel rows further on);
//   B (N side) = dZ, N = 64;  one accumulator per tap pair, all taps of the kernel in ONE unit, so dZ is loaded once per
//                pixel tile and X once as a 2-D halo box (TH x TW tiles, e.g. 4 x 32 -> 1.6x over-fetch for 3x3).
// Work unit = (source, 64-channel chunk of X, tap set sharing one halo box) x pixel split; epilogue lane = (tap, ci).
constexpr int kWgtMaxTaps = 16;

struct WgtSet {
  int16_t c_mul; int8_t p; int8_t dh0, dw0;
  uint8_t ntaps;
  uint8_t ddh[kWgtMaxTaps], ddw[kWgtMaxTaps], kidx[kWgtMaxTaps];
};

struct WgtK {
  int n, grid_h, grid_w;
  int TW, TH, tiles_w, tiles_h, tw_shift;
  int total_tiles, splits;
  int nsets;
  WgtSet sets[4];
  int chunks0, chunks1, c0, c1, pitch0, pitch1;
  int halo_w, halo_h;
  int N, s_groups;                 // dZ channels per MMA (multiple of 16, <= 256) and the 64-channel boxes that hold them
  int nparts;                      // dZ channel ranges of N each (> 1 when all taps x all dZ channels do not fit TMEM, e.g. 192)
  int s_bytes, l_bytes, stage_bytes, stages;
  int tmem_cols;
  uint32_t idesc;
  int cs, ktot, m_pad;
  float* ws;
  int* err_flag;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_t_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmL0,
                    const __grid_constant__ CUtensorMap tmL1, const __grid_constant__ WgtK P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + (uint32_t)P.stages * (uint32_t)P.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgMaxStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgMaxStages);
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(base_ptr + (size_t)P.stages * P.stage_bytes + 8u * (2 * kWgMaxStages + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // blockIdx.x = unit * splits + split ; unit = (set * nchunks + chunk) * nparts + part
  const int split = (int)(blockIdx.x % (unsigned)P.splits);
  int unit = (int)(blockIdx.x / (unsigned)P.splits);
  const int n_base = (unit % P.nparts) * P.N;     // first dZ channel of this unit
  unit /= P.nparts;
  const int nchunks = P.chunks0 + P.chunks1;
  const int chunk = unit % nchunks;
  const WgtSet S = P.sets[unit / nchunks];
  const bool src1 = chunk >= P.chunks0;
  const int ch_base = (src1 ? chunk - P.chunks0 : chunk) * 64;
  const int c_src = src1 ? P.c1 : P.c0;
  const int pitch = src1 ? P.pitch1 : P.pitch0;
  const CUtensorMap* tmL = src1 ? &tmL1 : &tmL0;
  const int mtiles = (S.ntaps + 1) >> 1;
  const int per = (P.total_tiles + P.splits - 1) / P.splits;
  const int t_begin = min(P.total_tiles, split * per), t_end = min(P.total_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(tmL);
    for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================================================= producer: one dZ box + one X halo box per pixel tile
    int slot = 0; uint32_t phase = 0;
    const uint32_t tx = (uint32_t)P.s_groups * (uint32_t)(128 * 128) + (uint32_t)(P.halo_w * P.halo_h * 128);
    for (int t = t_begin; t < t_end; ++t) {
      int q = t;
      const int tw = q % P.tiles_w; q /= P.tiles_w;
      const int th = q % P.tiles_h;
      const int img = q / P.tiles_h;
      const int w0 = tw * P.TW, h0 = th * P.TH;
      mbar_wait(empty_bar(slot), phase ^ 1u, P.err_flag, 21);
      if (elect_one()) {
        const uint32_t sbase = base + (uint32_t)slot * (uint32_t)P.stage_bytes;
        mbar_expect_tx(full_bar(slot), tx);
        for (int g = 0; g < P.s_groups; ++g)
          tma_load_5d(sbase + (uint32_t)g * (128u * 128u), &tmS, full_bar(slot), n_base + g * 64, w0, 0, h0, img);
        tma_load_5d(sbase + (uint32_t)P.s_bytes, tmL, full_bar(slot), S.c_mul * pitch + ch_base, w0 + S.dw0, S.p, h0 + S.dh0, img);
      }
      __syncwarp();
      if (++slot == P.stages) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    int slot = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    const uint64_t b_hi = make_mnmajor_desc(0, 128u * 128u, 1024u);    // dZ: 64-channel groups one box apart
    // per M tile: pixel-row offset of its first tap and the distance to its second tap (both in 16-byte units)
    uint32_t a_off[kWgtMaxTaps / 2];
    uint64_t a_hi[kWgtMaxTaps / 2];
#pragma unroll
    for (int i = 0; i < kWgtMaxTaps / 2; ++i) {
      const int ja = min(2 * i, (int)S.ntaps - 1), jb = min(2 * i + 1, (int)S.ntaps - 1);
      const int oa = S.ddh[ja] * P.halo_w + S.ddw[ja], ob = S.ddh[jb] * P.halo_w + S.ddw[jb];
      a_off[i] = (uint32_t)oa * 8u;
      a_hi[i] = make_mnmajor_desc(0, (uint32_t)(ob - oa) * 128u, 1024u);
    }
    uint32_t k_off[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int px = kk * 16;
      k_off[kk] = (uint32_t)(((px >> P.tw_shift) * P.halo_w + (px & (P.TW - 1))) * 8);
    }
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full_bar(slot), phase, P.err_flag, 22);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sbase = base + (uint32_t)slot * (uint32_t)P.stage_bytes;
        const uint32_t b_lo = (sbase & 0x3FFFFu) >> 4;
        const uint32_t a_lo = ((sbase + (uint32_t)P.s_bytes) & 0x3FFFFu) >> 4;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t b = b_hi | (uint64_t)(b_lo + (uint32_t)kk * 128u);
#pragma unroll
          for (int i = 0; i < kWgtMaxTaps / 2; ++i) {
            if (i < mtiles) {
              const uint64_t a = a_hi[i] | (uint64_t)(a_lo + k_off[kk] + a_off[i]);
              umma_bf16(tmem_base + (uint32_t)(i * P.N), a, b, P.idesc, accumulate | (uint32_t)kk);
            }
          }
        }
        umma_commit(empty_bar(slot));
        if (t == t_end - 1) umma_commit(done_bar);
      }
      __syncwarp();
      accumulate = 1;
      if (++slot == P.stages) { slot = 0; phase ^= 1u; }
    }
  } else {
    // ======================================================= epilogue: lane = (tap of the pair, X channel); columns = dZ channels
    const int q = warp & 3;
    const int l = q * 32 + lane;
    const int ci = l & 63;
    const int ctot = P.c0 + P.c1;
    const int k_src = src1 ? P.c0 : 0;
    const bool have = t_end > t_begin;
    if (have) {
      mbar_wait(done_bar, 0, P.err_flag, 23);
      tc_fence_after();
    }
    float* slab = P.ws + (size_t)split * P.m_pad * P.ktot;
    for (int i = 0; i < mtiles; ++i) {
      const int j = 2 * i + (l >> 6);
      const bool row_ok = j < S.ntaps && ch_base + ci < c_src;
      const size_t kcol = row_ok ? (size_t)S.kidx[j] * ctot + k_src + ch_base + ci : 0;
      for (int c16 = 0; c16 < P.N / 16; ++c16) {
        if (n_base + c16 * 16 >= P.cs) break;
        float v[16];
        if (have) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * P.N + c16 * 16), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0.f;
        }
        if (row_ok) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int co = n_base + c16 * 16 + k;
            if (co < P.cs) slab[(size_t)co * P.ktot + kcol] = v[k];     // consecutive lanes = consecutive ci: coalesced
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

struct WgtPlan {
  WgtK P;
  int units, smem;
  size_t ws_bytes;
};

int wgt_plan(const adb_wgrad_desc* d, WgtPlan& pl, int sm_count, int max_smem) {
  WgtK& P = pl.P;
  memset(&P, 0, sizeof(P));
  struct Raw { int c_mul, p, dw, dh, kidx; };
  Raw raw[64];
  const int ntaps = d->kh * d->kw;
  if (d->kind == ADB_CONV_S1) {
    const int pad_h = (d->kh - 1) / 2, pad_w = (d->kw - 1) / 2;
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) raw[r * d->kw + s] = {0, 0, s - pad_w, r - pad_h, r * d->kw + s};
    P.grid_h = d->h_in; P.grid_w = d->w_in;
  } else {
    auto fl2 = [](int u) { return (u >= 0) ? u / 2 : -((-u + 1) / 2); };
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) {
        const int u = r - d->pad, v = s - d->pad;
        raw[r * d->kw + s] = {v - 2 * fl2(v), u - 2 * fl2(u), fl2(v), fl2(u), r * d->kw + s};
      }
    P.grid_h = d->h_in / 2; P.grid_w = d->w_in / 2;
  }
  // tap sets: same (c_mul, p) -> one halo box; taps ordered by (dh, dw) so that pair offsets are non-negative
  bool used[64] = {false};
  int ew = 0, eh = 0;
  P.nsets = 0;
  for (int t0 = 0; t0 < ntaps; ++t0) {
    if (used[t0]) continue;
    if (P.nsets >= 4) return 1;
    WgtSet& s = P.sets[P.nsets++];
    int members[64], nm = 0, dw_min = 1 << 20, dw_max = -(1 << 20), dh_min = 1 << 20, dh_max = -(1 << 20);
    for (int t1 = t0; t1 < ntaps; ++t1) {
      if (used[t1] || raw[t1].c_mul != raw[t0].c_mul || raw[t1].p != raw[t0].p) continue;
      members[nm++] = t1; used[t1] = true;
      dw_min = std::min(dw_min, raw[t1].dw); dw_max = std::max(dw_max, raw[t1].dw);
      dh_min = std::min(dh_min, raw[t1].dh); dh_max = std::max(dh_max, raw[t1].dh);
    }
    if (nm > kWgtMaxTaps) return 1;
    std::sort(members, members + nm, [&](int a, int b) { return raw[a].dh != raw[b].dh ? raw[a].dh < raw[b].dh : raw[a].dw < raw[b].dw; });
    s.c_mul = (int16_t)raw[t0].c_mul; s.p = (int8_t)raw[t0].p; s.dh0 = (int8_t)dh_min; s.dw0 = (int8_t)dw_min; s.ntaps = (uint8_t)nm;
    for (int k = 0; k < nm; ++k) {
      s.ddh[k] = (uint8_t)(raw[members[k]].dh - dh_min); s.ddw[k] = (uint8_t)(raw[members[k]].dw - dw_min);
      s.kidx[k] = (uint8_t)raw[members[k]].kidx;
    }
    ew = std::max(ew, dw_max - dw_min); eh = std::max(eh, dh_max - dh_min);
  }
  // pixel tile: 128 = TH x TW with TW >= 16; prefer 4 x 32 (small halo) when the map is wide enough
  int TW = 32;
  while (TW > P.grid_w && TW > 16) TW >>= 1;
  P.TW = TW; P.TH = 128 / TW;
  P.tw_shift = 0;
  while ((1 << P.tw_shift) < TW) ++P.tw_shift;
  P.tiles_w = (P.grid_w + TW - 1) / TW;
  P.tiles_h = (P.grid_h + P.TH - 1) / P.TH;
  const long long tt = (long long)d->n * P.tiles_w * P.tiles_h;
  if (tt >= (1LL << 30)) return 1;
  P.total_tiles = (int)tt;
  P.n = d->n;
  P.halo_w = TW + ew; P.halo_h = P.TH + eh;
  if (P.halo_w > 256 || P.halo_h > 256) return 1;
  P.c0 = d->c0; P.c1 = d->c1; P.pitch0 = d->c0_pitch; P.pitch1 = d->act1 ? d->c1_pitch : d->c0_pitch;
  P.chunks0 = (d->c0 + 63) / 64; P.chunks1 = (d->c1 + 63) / 64;
  int max_tiles = 1;
  for (int i = 0; i < P.nsets; ++i) max_tiles = std::max(max_tiles, (P.sets[i].ntaps + 1) / 2);
  // one accumulator per tap pair must fit TMEM (512 columns): gradients with an odd multiple of 64 channels that do not
  // fit (192 x 5 tap pairs) are cut into equal dZ channel ranges, one unit each (X is then read once per range)
  P.nparts = 1;
  P.N = (d->cg + 15) / 16 * 16;
  if (d->cg % 128 != 0)
    while (P.nparts < 4 && (P.N > 256 || max_tiles * P.N > 512)) {
      ++P.nparts;
      P.N = ((d->cg + P.nparts - 1) / P.nparts + 15) / 16 * 16;
    }
  if (P.N > 256) return 1;
  P.s_groups = (P.N + 63) / 64;
  int cols = 32;
  while (cols < max_tiles * P.N) cols <<= 1;
  if (cols > 512) return 1;
  P.tmem_cols = cols;
  P.idesc = make_idesc_bf16(128u, (uint32_t)P.N) | (1u << 15) | (1u << 16);
  P.cs = d->cg;
  P.m_pad = (d->cg + 15) / 16 * 16;
  P.ktot = ntaps * (d->c0 + d->c1);
  P.s_bytes = P.s_groups * 128 * 128;
  // the last tap pair's second view may start up to (off_b) rows in and runs 128 rows: keep the slack inside the slot
  P.l_bytes = wg_round_up((P.halo_w * P.halo_h + 16) * 128, 1024);
  P.stage_bytes = P.s_bytes + P.l_bytes;
  const int bar_bytes = 8 * (2 * kWgMaxStages + 1) + 16;
  int stages = std::min((max_smem - 1024 - bar_bytes) / P.stage_bytes, kWgMaxStages);
  if (stages < 2) return 1;
  P.stages = stages;
  pl.smem = stages * P.stage_bytes + bar_bytes + 1024;
  pl.units = P.nsets * (P.chunks0 + P.chunks1) * P.nparts;
  int splits = std::max(1, (2 * sm_count + pl.units - 1) / pl.units);
  splits = std::min(splits, std::max(1, P.total_tiles / 4));
  splits = std::min(splits, P.total_tiles);
  const size_t slab = (size_t)P.m_pad * P.ktot * sizeof(float);
  if (d->workspace_bytes > 0) splits = (int)std::max<long long>(0, std::min<long long>(splits, (long long)(d->workspace_bytes / (long long)slab)));
  P.splits = splits;
  pl.ws_bytes = (size_t)std::max(splits, 1) * slab;
  return 0;
}

// The tap-packed kernel is chosen when it fits (one TMEM accumulator per tap pair) and the channel-major kernel would run
// part-empty M tiles (gradient channels not a multiple of 128) — measured 1.2-2.3x faster there (profiles/r1_train_summary.md).
inline bool wgt_wanted(const adb_wgrad_desc* d) {
  return d->mode == 2 || (d->mode == 0 && (d->cg % 128 != 0 || d->kind == ADB_CONV_S2));
}

// 5-D view {C, W, P, H, N} of an NHWC bf16 buffer (same convention as conv_igemm.cu)
int wg_act_tmap(CUtensorMap* m, const void* base, int c_dim, int pitch, int n, int h, int w, bool s2d, int box_w, int box_h) {
  uint64_t dims[5], strides[4];
  uint32_t box[5] = {64u, (uint32_t)box_w, 1u, (uint32_t)box_h, 1u};
  const uint64_t px = (uint64_t)pitch * 2;
  if (!s2d) {
    dims[0] = c_dim; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = n;
    strides[0] = px; strides[1] = px * w; strides[2] = px * w; strides[3] = px * w * h;
  } else {
    dims[0] = 2 * (uint64_t)pitch; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = n;
    strides[0] = 2 * px; strides[1] = px * w; strides[2] = 2 * px * w; strides[3] = px * w * h;
  }
  return adbh::make_tmap_bf16(m, base, 5, dims, strides, box, 128);
}

}  // namespace

extern "C" int64_t adb_wgrad_workspace_bytes(const adb_wgrad_desc* d) {
  WgPlan pl;
  adbh::DeviceInfo di;
  int sm = 148, smem = 227 * 1024;
  if (adbh::device_info(&di) == ADB_OK) { sm = di.sm_count; smem = di.max_smem_optin; }
  adb_wgrad_desc q = *d;
  q.workspace_bytes = 0;
  if (wg_plan(&q, pl, sm, smem) != ADB_OK) return -1;
  if (wgt_wanted(&q)) {
    WgtPlan tp;
    if (wgt_plan(&q, tp, sm, smem) == 0) return (int64_t)std::max(tp.ws_bytes, pl.ws_bytes);
  }
  return (int64_t)pl.ws_bytes;
}

extern "C" int adb_wgrad(const adb_wgrad_desc* d, void* stream) {
  adbh::DeviceInfo di;
  int st = adbh::device_info(&di);
  if (st != ADB_OK) return st;
  if (di.cc_major != 10) return adbh::fail(ADB_ERR_NO_DEVICE, "adb_wgrad: device sm_%d%d is not sm_100", di.cc_major, di.cc_minor);
  WgPlan pl;
  st = wg_plan(d, pl, di.sm_count, di.max_smem_optin);
  if (st != ADB_OK) return st;
  WgK& P = pl.P;
  ADB_REQUIRE(d->workspace && d->dw, "adb_wgrad: null workspace / dw");
  ADB_REQUIRE(P.splits >= 1, "adb_wgrad: workspace of %lld bytes cannot hold one %d x %d fp32 slab", (long long)d->workspace_bytes, P.m_pad, P.ktot);
  ADB_REQUIRE(d->layout == ADB_WG_OIHW || (d->layout == ADB_WG_STEM && d->kw == 1 && d->act1 == nullptr && d->stem_kw >= 1 && 3 * d->stem_kw <= d->c0),
              "adb_wgrad: bad output layout %d", d->layout);
  P.ws = d->workspace;
  P.err_flag = adbh::kernel_err_flag();

  alignas(64) CUtensorMap tmS, tmL0, tmL1;
  WgtPlan tp;
  if (wgt_wanted(d) && wgt_plan(d, tp, di.sm_count, di.max_smem_optin) == 0 && tp.P.splits >= 1) {
    WgtK& T = tp.P;
    T.ws = d->workspace;
    T.err_flag = P.err_flag;
    const bool s2 = d->kind == ADB_CONV_S2;
    st = wg_act_tmap(&tmS, d->grad, d->cg, d->cg_pitch, d->n, T.grid_h, T.grid_w, false, T.TW, T.TH);
    if (st != ADB_OK) return st;
    st = wg_act_tmap(&tmL0, d->act0, d->c0_pitch, d->c0_pitch, d->n, d->h_in, d->w_in, s2, T.halo_w, T.halo_h);
    if (st != ADB_OK) return st;
    if (d->act1) {
      st = wg_act_tmap(&tmL1, d->act1, d->c1_pitch, d->c1_pitch, d->n, d->h_in, d->w_in, s2, T.halo_w, T.halo_h);
      if (st != ADB_OK) return st;
    } else {
      tmL1 = tmL0;
    }
    static bool configured_t = false;
    if (!configured_t) {
      ADB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
      configured_t = true;
    }
    const int smem_t = std::max(tp.smem, 120 * 1024);
    conv_wgrad_t_kernel<<<(unsigned)((long long)tp.units * T.splits), kWgThreads, smem_t, (cudaStream_t)stream>>>(tmS, tmL0, tmL1, T);
    ADB_CUDA_OK(cudaGetLastError());
    const long long total_t = (long long)T.cs * T.ktot;
    const int rgrid_t = (int)std::max<long long>(1, std::min<long long>((total_t + 255) / 256, (long long)di.sm_count * 8));
    wgrad_reduce_kernel<<<rgrid_t, 256, 0, (cudaStream_t)stream>>>(T.ws, T.splits, T.m_pad, (d->cg_true > 0 ? d->cg_true : d->cg), T.ktot,
                                                                    d->c0 + d->c1, d->layout == ADB_WG_STEM ? d->kh : d->kh * d->kw, d->layout,
                                                                    d->stem_kw, d->accumulate, d->dw);
    ADB_CUDA_OK(cudaGetLastError());
    return ADB_OK;
  }
  // channel extent = cg, not the pitch: `grad` may be a channel slice of a wider buffer (DenseNet block-buffer gradient) and a
  // 64-channel box that runs past the slice must be zero-filled by TMA rather than read past the end of the allocation
  st = wg_act_tmap(&tmS, d->grad, d->cg, d->cg_pitch, d->n, P.grid_h, P.grid_w, false, P.TW, P.TH);
  if (st != ADB_OK) return st;
  const bool s2d = d->kind == ADB_CONV_S2;
  st = wg_act_tmap(&tmL0, d->act0, d->c0_pitch, d->c0_pitch, d->n, d->h_in, d->w_in, s2d, pl.box_w, pl.box_h);
  if (st != ADB_OK) return st;
  if (d->act1) {
    st = wg_act_tmap(&tmL1, d->act1, d->c1_pitch, d->c1_pitch, d->n, d->h_in, d->w_in, s2d, pl.box_w, pl.box_h);
    if (st != ADB_OK) return st;
  } else {
    tmL1 = tmL0;
  }
  static bool configured = false;
  if (!configured) {
    ADB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
    configured = true;
  }
  const int smem = std::max(pl.smem, 120 * 1024);   // one CTA per SM: the CTA owns the SM's TMEM
  const long long grid = (long long)pl.units * P.splits;
  ADB_REQUIRE(grid < (1LL << 31), "adb_wgrad: grid too large");
  conv_wgrad_kernel<<<(unsigned)grid, kWgThreads, smem, (cudaStream_t)stream>>>(tmS, tmL0, tmL1, P);
  ADB_CUDA_OK(cudaGetLastError());
  const long long total = (long long)P.cs * P.ktot;
  const int rgrid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)di.sm_count * 8));
  const int ctot = d->c0 + d->c1;
  wgrad_reduce_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(P.ws, P.splits, P.m_pad, (d->cg_true > 0 ? d->cg_true : d->cg), P.ktot, ctot,
                                                                  d->layout == ADB_WG_STEM ? d->kh : d->kh * d->kw, d->layout,
                                                                  d->stem_kw, d->accumulate, d->dw);
  ADB_CUDA_OK(cudaGetLastError());
  return ADB_OK;
}

extern "C" double adb_wgrad_flops(const adb_wgrad_desc* d) {
  if (!d) return 0.0;
  const double oh = d->kind == ADB_CONV_S2 ? d->h_in / 2 : d->h_in, ow = d->kind == ADB_CONV_S2 ? d->w_in / 2 : d->w_in;
  return 2.0 * d->n * oh * ow * d->kh * d->kw * ((double)d->c0 + d->c1) * (d->cg_true > 0 ? d->cg_true : d->cg);
}
