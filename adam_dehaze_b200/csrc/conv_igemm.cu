// conv_igemm.cu — implicit-GEMM convolution for sm_100a: TMA-fed, tcgen05.mma into TMEM, fused epilogues.
//
// GEMM view:  D[pixel, cout] = sum_k A[pixel, k] * B[cout, k],  k = (tap, input channel).
//   A: NHWC bf16 feature map(s).  One "A load" = one chunk of Ck channels of the tile's input HALO: a 5-D TMA box
//      {Ck, TW+ew, 1, TH*MT+eh, 1} fetched ONCE and then read by every filter tap that falls inside it through a
//      row-shifted UMMA shared-memory descriptor (a 3x3 conv re-uses each halo box for 9 taps instead of re-reading the
//      input 9x through L2, which is what bounds an implicit GEMM on B200: L2->SM bandwidth for non-replicated data is
//      ~6.7 TB/s, about HBM speed).  Out-of-image pixels are zero-filled by the TMA unit = the conv's zero padding.
//      Stride-2 convs read the buffer through a space-to-depth view {2C, W/2, 2, H/2, N} (one A load per row/column
//      phase); ConvTranspose2d(4,2,1) runs as four 2x2 sub-pixel phases whose outputs are stored through the same view
//      of the destination.  A channel concat is two tensor maps walked back to back.  When the tile is not one image
//      row wide (TW < 128) and taps shift columns, every tap becomes its own A load (no halo re-use).
//   B: packed weights [cout][k] bf16, K-major, one 2-D TMA box {Ck, BN} per (tap, chunk), on its own mbarrier ring and
//      its own producer warp.
//   D: fp32 accumulators in TMEM, 128 pixels (lanes) x BN columns; MT sub-tiles share each B box.
// Roles (256 threads): warp 0 = A (halo) TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 = B (weights)
// TMA producer, warps 4-7 = epilogue (TMEM -> registers -> affine/residual/activation -> swizzled smem -> TMA store, or the fused
// image/dot epilogues).  The grid is persistent: one CTA per SM walking tiles round-robin; TMEM accumulators are
// double buffered when they fit so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Reference arithmetic replaced: nn.Conv2d/ConvTranspose2d + BatchNorm2d(eval) + ReLU/Tanh/Sigmoid (+ residual add),
// /root/reference/models/dehazing/base_model.py:4-41 and the branch forwards low_intensity.py:33-45,
// medium_intensity.py:78-117, high_intensity.py:92-138.
#include "adb_ptx.cuh"
#include "adb_host.h"
#include "conv_common.cuh"
#include <algorithm>
#include <math.h>

namespace {

using namespace adb;
using namespace adbc;

constexpr int kThreads = 384;         // 4 role warps + 8 epilogue warps
constexpr int kThreadsPre = 512;      // + 4 operand-transform warps (pre-activation fused into the A operand, DenseNet)
constexpr int kEpiWarp0 = 4;          // first epilogue warp
constexpr int kEpiWarps = 8;
constexpr int kPreWarp0 = kEpiWarp0 + kEpiWarps;   // first transform warp (kPre kernels only)
constexpr int kMaxTaps = 25;        // up to 5x5 (AlexNet conv2 inside LPIPS, loss.py:91)
constexpr int kMaxGroups = 4;
constexpr int kMaxStages = 12;

constexpr int kMaxALoads = 25;
constexpr int kMaxASlots = 8;
constexpr int kMaxBSlots = 12;

struct ALoad {          // one halo box per (group, phase): coordinates relative to the tile origin
  int16_t c_mul;        // channel coordinate = c_mul * channel_pitch(src) + chunk*Ck   (space-to-depth column phase)
  int8_t p;             // row-phase coordinate
  int8_t dw0, dh0;      // box origin = (w0 + dw0, h0 + dh0)
  uint8_t tap_begin, tap_count;
};
struct TapK {
  uint16_t shift_px;    // pixel offset of this tap's 128-row view inside the halo box
  uint16_t kidx;        // position of the tap in the packed-weight K order (r*kw + s)
};

struct ConvK {
  // batch
  int n, n_start;
  const int* n_dev;
  // tile grid (per group)
  int grid_h, grid_w;      // extent of the pixel grid tiles cover (output H/W; input H/W for ConvTranspose phases)
  int TW, TH, MT;
  int tiles_w, tiles_h;
  FastDiv fd_nt, fd_g, fd_tw, fd_th;   // decode_tile divisors: n_tiles_n, ngroups, tiles_w, tiles_h
  int tw_shift;            // log2(TW)
  int ncta;                // 1, or 2 = CTA-pair mode (cluster of 2, tcgen05 cta_group::2): a tile spans both CTAs' sub-tiles
  // K walk
  int Ck, row_bytes;
  int chunks0, chunks1, pitch0, pitch1, ctot;
  int ngroups;
  int n_aloads[kMaxGroups];
  ALoad aloads[kMaxGroups][kMaxALoads];
  TapK taps[kMaxGroups][kMaxTaps];
  int ntaps;               // taps per group (all groups equal)
  int halo_w;              // TW + ew: pixels per halo row
  int sub_px;              // pixel offset between the MT sub-tiles inside the halo box (= TH * halo_w)
  // N
  int BN, n_tiles_n, bn_cols, cout_pad;
  // pipeline
  int a_slots, b_slots, acc_stages, tmem_cols;
  int a_slot_bytes, b_slot_bytes;     // smem slot sizes (1024-aligned)
  int a_tx_bytes, b_tx_bytes;         // bytes one TMA box delivers
  int taps_per_slot, b_tap_stride;    // a weight slot carries up to taps_per_slot tap boxes, b_tap_stride bytes apart
  uint32_t idesc;
  int desc_base_offset;               // experiment knob: put (addr>>7)&7 in the descriptor's base_offset field
  // epilogue
  int epi, act;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int res_pitch;
  int Cs, n_slabs, slab_bytes, n_slab_bufs;   // FEATURE: output slab channels (TMA store box inner dim), staging ring
  int out_c_off[kMaxGroups];             // 5-D store coordinate 0 base per group
  int out_p[kMaxGroups];                 // 5-D store coordinate 2 per group
  const float* dot_w; float dot_b; float* dot_out;
  int img_mode;
  const float* img_x; float* img_out; const int* img_index; const float* img_guidance; const float* img_alpha;
  int c0;                             // channels of src0 (weight K offset of src1's first chunk)
  int ks_last0, ks_last1;             // K steps (of 16 channels) in the last chunk of src0 / src1 (ragged 64-channel chunking)
  int a_sbo_bytes;                    // byte distance between the 8-row groups of an A view (8*row_bytes, or one halo row for 8-wide 2-D tiles)
  const float* pre_scale;             // optional pre-activation relu(x*pre_scale[c] + pre_shift[c]) applied to the A operand
  const float* pre_shift;
  float* stat_out;                    // optional per-(32-row group, channel) partials of the stored output (adb_conv_desc.stat_out)
  int stat_mode;                      // 1: (sum, sum of squares)   2: (sum, max)
  int* err_flag;
  long long* dbg;   // optional timeline buffer (tune_flags bit 2): [6 roles][256 events] of clock64() from CTA 0
  int dbg_detail;   // tune_flags bit 6: the buffer instead holds a flat sequence of epilogue sub-step stamps (warp 4, lane 0)
};

struct SmemLayout {
  // byte offsets from the 1024-aligned base
  uint32_t a_off, b_off, slab_off, scale_off, pre_off, bar_off, tap_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int a_slots, int a_slot_bytes, int b_slots, int b_slot_bytes,
                                                  int slab_bytes, int n_slab_bufs, int cout_pad, int pre_channels = 0) {
  SmemLayout L;
  uint32_t off = 0;
  L.a_off = off; off += (uint32_t)a_slots * a_slot_bytes;
  L.b_off = off; off += (uint32_t)b_slots * b_slot_bytes;
  L.slab_off = off; off += (uint32_t)n_slab_bufs * slab_bytes;
  L.scale_off = off; off += (uint32_t)cout_pad * 8;      // scale then shift, fp32
  off = (off + 15u) & ~15u;
  L.pre_off = off; off += (uint32_t)pre_channels * 8;    // pre-activation scale then shift, fp32 (kPre kernels)
  off = (off + 15u) & ~15u;
  L.bar_off = off; off += 8u * (2 * kMaxASlots + 2 * kMaxBSlots + 4) + 16 + 8u * kMaxASlots;  // fullA, emptyA, fullB, emptyB, tmem full/empty, tmem ptr, readyA
  L.tap_off = off; off += 4u * kMaxGroups * kMaxTaps;    // per-tap A descriptor offsets (16-byte units) for the MMA issuer
  L.total = off;
  return L;
}

#define ADB_DBG(role, idx) do { if (P.dbg && !P.dbg_detail && blockIdx.x == 0 && lane == 0 && (idx) < 256) P.dbg[(role) * 256 + (idx)] = clock64(); } while (0)
// epilogue sub-step stamp: tag in the top byte, clock below (flat sequence, warp 4 lane 0 of CTA 0)
#define ADB_DBGE(tag) do { if (P.dbg && P.dbg_detail && blockIdx.x == 0 && ew == 0 && half == 0 && lane == 0 && e_i < 6 * 256) P.dbg[e_i++] = ((long long)(tag) << 56) | (clock64() & 0x00FFFFFFFFFFFFFFLL); } while (0)

// D[tmem] (+)= A[smem] * B[smem]; descriptors given as (low word, high word) so that advancing a start address is one 32-bit add.
template <bool kPair>
__device__ __forceinline__ void umma_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  if (kPair) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

struct IssueK {          // per-launch constants of the MMA issuer
  uint32_t a_hi, b_hi;   // descriptor high words (SBO, version, swizzle) of the A views / weight boxes
  uint32_t sub_step;     // A start-address step between the MT sub-tiles
  uint32_t b_tap_step;   // B start-address step between the tap boxes of one weight slot
  uint32_t idesc;
};

// The MMAs of one weight slot: `nt` taps (A view = slot + tap offset from the shared-memory table), kKs K steps of 16
// (start address += 32 B each) for kMt sub-tiles sharing each B box.  A few integer instructions per MMA: every MMA of a
// narrow tile takes only ~N/2 cycles, so the issue path must not cost more (measured: 250-350 cycles of descriptor
// arithmetic per tap in the first version kept N <= 96 layers at 0.6 of the tensor peak).
template <bool kPair, int kMt, int kKs>
__device__ __forceinline__ void issue_taps(const IssueK& K, uint32_t d0, uint32_t d1, uint32_t a_lo, uint32_t b_lo, const uint32_t* tap_off,
                                           int nt, uint32_t accumulate) {
#pragma unroll 1
  for (int jj = 0; jj < nt; ++jj) {
    const uint32_t a = a_lo + tap_off[jj];
    const uint32_t b = b_lo + (uint32_t)jj * K.b_tap_step;
#pragma unroll
    for (int kk = 0; kk < kKs; ++kk) {
      const uint32_t acc = (kk == 0) ? (accumulate | (uint32_t)jj) : 1u;      // only the tile's very first MMA overwrites
      umma_lh<kPair>(d0, a + (uint32_t)(kk * 2), K.a_hi, b + (uint32_t)(kk * 2), K.b_hi, K.idesc, acc);
      if (kMt == 2) umma_lh<kPair>(d1, a + K.sub_step + (uint32_t)(kk * 2), K.a_hi, b + (uint32_t)(kk * 2), K.b_hi, K.idesc, acc);
    }
  }
}

struct TileCoord { int nt, g, w0, h0, img; };

// Epilogue building blocks.  An epilogue warp handles "slabs": 32 tile rows x CS16*16 channels.  The residual row of the
// NEXT slab is requested while the current one is computed, and all TMEM columns of a slab are requested before the
// single tcgen05.wait::ld, so a slab exposes neither a global-load latency nor one TMEM round trip per 16 channels.
// Per-tile addressing of this lane's residual fetches: load i of a slab covers tile row ew*32 + i*(32/CPR) + lane/CPR, 16-byte
// chunk lane%CPR (consecutive lanes read consecutive chunks of a row: coalesced).  With tiles at least 32 pixels wide (every
// map wider than 16 pixels) a warp's 32 rows lie in one image row, so load i sits i*(32/CPR) pixels after load 0: a base
// pointer and the count of in-image loads describe the lane; narrower tiles take the general form.
struct ResLane {
  const __nv_bfloat16* base;   // load 0 of item (sub-tile 0, slab 0)
  int nvalid;                  // loads 0 .. nvalid-1 are inside the image width (wide tiles)
  int dh0;                     // image-row offset of this warp's rows inside a sub-tile (wide tiles)
  uint32_t step;               // elements between consecutive loads of an item (wide tiles)
  uint32_t mt_step;            // elements between the sub-tiles of a tile
};
template <int CS16>
__device__ __forceinline__ ResLane res_lane(const ConvK& P, const TileCoord& tc, int ew, int lane) {
  constexpr int CPR = CS16 * 2, RS = 32 / CPR;
  ResLane r;
  const int row0 = ew * 32 + lane / CPR;
  r.dh0 = row0 >> P.tw_shift;
  const int w = tc.w0 + (row0 & (P.TW - 1));
  r.base = P.residual + (((size_t)tc.img * P.grid_h + tc.h0 + r.dh0) * (size_t)P.grid_w + w) * P.res_pitch + tc.nt * P.BN + (lane % CPR) * 8;
  r.nvalid = min(CPR, max(0, (P.grid_w - w + RS - 1) / RS));
  r.step = (uint32_t)(RS * P.res_pitch);
  r.mt_step = (uint32_t)(P.TH * P.grid_w * P.res_pitch);     // < 2^31: at most 16 image rows of one map
  // kept in registers (or a cheap local-memory slot): left alone, the compiler re-derives all of it from the tile index in
  // front of every slab's loads — a 500-cycle chain of 64-bit multiplies and constant-bank reads on the epilogue's critical path
  asm volatile("" : "+l"(r.base), "+r"(r.nvalid), "+r"(r.dh0), "+r"(r.step), "+r"(r.mt_step));
  return r;
}
// (mt, sl) = the item's sub-tile and slab.  kPrefetch: only pull the rows into L2 (prefetch.global.L2; no registers held) — the
// later register load then pays an L2 round trip instead of a DRAM one.
// Loads [I0, I1) of the item only (the caller spreads an item's loads over its slab: eight warps firing all of theirs at once
// stall at issue for about a memory round trip).
template <int CS16, bool kPrefetch = false, int I0 = 0, int I1 = CS16 * 2>
__device__ __forceinline__ void load_residual(const ConvK& P, const TileCoord& tc, const ResLane& rl, int mt, int sl, bool live,
                                              int ew, int lane, uint4 (&q)[CS16 * 2]) {
  constexpr int CPR = CS16 * 2, RS = 32 / CPR;
  constexpr int Cs = CS16 * 16;
  if (!live) return;
  const int h_left = P.grid_h - tc.h0 - mt * P.TH;       // image rows left from the sub-tile's first row on
  if (P.TW >= 32) {
    const __nv_bfloat16* p = rl.base + ((uint32_t)mt * rl.mt_step + (uint32_t)(sl * Cs));
    const int nv = rl.dh0 < h_left ? rl.nvalid : 0;
    if (kPrefetch) {
#pragma unroll
      for (int i = I0; i < I1; ++i)
        if (i < nv) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (uint32_t)i * rl.step));
      return;
    }
#pragma unroll
    for (int i = I0; i < I1; ++i) q[i] = i < nv ? __ldg(reinterpret_cast<const uint4*>(p + (uint32_t)i * rl.step)) : make_uint4(0, 0, 0, 0);
  } else if (!kPrefetch) {
    const int ch = tc.nt * P.BN + sl * Cs + (lane % CPR) * 8;
#pragma unroll
    for (int i = I0; i < I1; ++i) {
      const int row = ew * 32 + i * RS + lane / CPR;
      const int h = tc.h0 + mt * P.TH + (row >> P.tw_shift), w = tc.w0 + (row & (P.TW - 1));
      q[i] = (h < P.grid_h && w < P.grid_w)
                 ? __ldg(reinterpret_cast<const uint4*>(P.residual + (((size_t)tc.img * P.grid_h + h) * P.grid_w + w) * P.res_pitch + ch))
                 : make_uint4(0, 0, 0, 0);
    }
  }
}

// `rank` = this CTA's rank in its pair (0 when ncta == 1): a pair's tile is two row-adjacent sub-tiles, one per CTA.
__device__ __forceinline__ TileCoord decode_tile(const ConvK& P, int t, int rank) {
  TileCoord c;
  uint32_t q = (uint32_t)t, r;
  fast_divmod(q, P.fd_nt, q, r); c.nt = (int)r;
  fast_divmod(q, P.fd_g, q, r);  c.g = (int)r;
  fast_divmod(q, P.fd_tw, q, r); c.w0 = (int)r * P.TW;
  fast_divmod(q, P.fd_th, q, r); c.h0 = ((int)r * P.ncta + rank) * (P.TH * P.MT);
  c.img = (int)q;
  return c;
}

// Channel partials of one staged slab (32 tile rows x CS16*16 channels, bf16, swizzled — exactly what the TMA store is about to
// write): lane l sums channels 2l, 2l+1 down the 32 rows (one conflict-free 4-byte shared load per row: the 32 lanes read one
// whole staging row), skipping rows outside the image.  Fixed order -> deterministic.  Consumers: BatchNorm2d(train) statistics
// without a pass over z (stat_mode 1 -> adb_bn_finalize_stats); AttentionBlock's global average / max pool without a pass over
// its input (stat_mode 2 -> adb_attn_pool_from_stats).
template <int CS16>
__device__ __forceinline__ void slab_stats(const ConvK& P, uint32_t sbuf, int lane, uint32_t valid_rows, size_t slot, int ch) {
  constexpr uint32_t span = CS16 * 32;
  constexpr int pairs = CS16 * 8;
  if (lane < pairs) {
    const bool mx = P.stat_mode == 2;
    float s0 = 0.f, s1 = 0.f;
    float a0 = mx ? -INFINITY : 0.f, a1 = a0;
    // the loads below are plain (non-volatile) asm so that eight of them are in flight at a time; their base address is
    // produced by a volatile asm that cannot move above the staging stores / fence before it, and the global stores of the
    // sums cannot sink below the next memory-clobbering asm, so the loads stay between this warp's writes of the buffer
    uint32_t base;
    asm volatile("mov.u32 %0, %1;" : "=r"(base) : "r"(sbuf) : "memory");
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 8) {
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm("ld.shared.b32 %0, [%1];" : "=r"(w[u]) : "r"(base + swizzle_addr((uint32_t)(r0 + u) * span + (uint32_t)lane * 4u, span)));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if ((valid_rows >> (r0 + u)) & 1u) {
          const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[u]);
          const float v0 = __low2float(b2), v1 = __high2float(b2);
          s0 += v0; s1 += v1;
          if (mx) { a0 = fmaxf(a0, v0); a1 = fmaxf(a1, v1); }
          else { a0 = fmaf(v0, v0, a0); a1 = fmaf(v1, v1, a1); }
        }
      }
    }
    float* o = P.stat_out + slot * 2 * (size_t)P.cout_pad + ch + 2 * lane;
    *reinterpret_cast<float2*>(o) = make_float2(s0, s1);
    *reinterpret_cast<float2*>(o + P.cout_pad) = make_float2(a0, a1);
  }
  __syncwarp();
}

// This warp has read its share of an accumulator stage: order the TMEM reads before the arrival and hand the stage back
// (8 arrivals per CTA complete the phase; a pair's arrivals all go to the leader CTA's barrier).
template <bool kPair>
__device__ __forceinline__ void release_acc(uint32_t tempty, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) { if (kPair) mbar_arrive_cluster(tempty); else mbar_arrive_relaxed(tempty); }
}

// FEATURE epilogue of one tile for one epilogue warp.  The tile's work items are its (sub-tile, slab) pairs in order; the two
// warps that share a TMEM lane quarter (`half` 0 / 1) take alternate items.  Each warp owns one staging buffer and one
// TMA-store bulk group stream; no cross-warp barrier anywhere.  The accumulator read and (without a residual) the whole slab
// arithmetic run BEFORE the wait for the previous TMA store to have read the staging buffer, so that wait hides behind them.
// The accumulator stage goes back to the MMA issuer (`tempty`) as soon as this warp's LAST slab is in registers — the rest of that
// slab's work (arithmetic, staging, store) no longer touches TMEM and overlaps the next tile's first MMAs.
template <int kAct, int CS16, bool kRes, bool kPair>
__device__ __forceinline__ void feature_tile(const ConvK& P, const CUtensorMap* tmOut, const TileCoord& tc, uint32_t tfull,
                                             uint32_t tfull_phase, uint32_t tempty, uint32_t tmem_tile, uint32_t sbuf, const float* s_scale,
                                             const float* s_shift, int ew, int half, int lane, int& e_i, size_t slot_base) {
  constexpr int Cs = CS16 * 16;
  const int items = P.MT * P.n_slabs;
  const int ch0 = tc.nt * P.BN;                        // first output channel of this N tile
  const int q_row0 = ew * 32;                          // first tile row of this warp
  const int q_th = q_row0 >> P.tw_shift, q_tw = q_row0 & (P.TW - 1);
  // TMA-store coordinates of this warp's rows: fixed for the tile (pinned so that they are not re-derived from the tile index
  // on the one lane that issues every store)
  int st_c = P.out_c_off[tc.g] + ch0, st_w = tc.w0 + q_tw, st_p = P.out_p[tc.g], st_h = tc.h0 + q_th, st_n = tc.img;
  asm volatile("" : "+r"(st_c), "+r"(st_w), "+r"(st_p), "+r"(st_h), "+r"(st_n));
  uint4 q[CS16 * 2], qn[CS16 * 2];
#pragma unroll
  for (int i = 0; i < CS16 * 2; ++i) q[i] = qn[i] = make_uint4(0, 0, 0, 0);
  ADB_DBGE(1);
  ResLane rl;
  // items advance by 2 per warp: (mt, sl) of the current item and (mtn, sln) of the one whose residual is in flight, stepped
  // without a division
  int mt = 0, sl = half;
  while (sl >= P.n_slabs) { sl -= P.n_slabs; ++mt; }
  int mtn = mt, sln = sl;
  if (kRes) {
    // Independent of the accumulator, so all of it overlaps this tile's main loop: the first item's rows go to registers, the
    // later items' rows are pulled into L2 (their register loads are issued one item ahead, which covers an L2 round trip but
    // not a DRAM one).
    rl = res_lane<CS16>(P, tc, ew, lane);
    load_residual<CS16>(P, tc, rl, mtn, sln, half < items, ew, lane, qn);
    int mtp = mt, slp = sl;
    for (int jp = half + 2; jp < items; jp += 2) {
      slp += 2;
      while (slp >= P.n_slabs) { slp -= P.n_slabs; ++mtp; }
      load_residual<CS16, true>(P, tc, rl, mtp, slp, true, ew, lane, qn);
    }
  }
  mbar_wait(tfull, tfull_phase, P.err_flag, 4);
  tc_fence_after();
  ADB_DBGE(2);
#pragma unroll 1
  for (int j = half; j < items; j += 2) {
    const int cl = sl * Cs;                            // channel offset of the slab inside the N tile
    const bool last = j + 2 >= items;
    float v[Cs];
    slab_tmem_load<CS16>(tmem_tile + (uint32_t)(mt * P.bn_cols + cl), v);
    uint32_t pk[CS16 * 8];
    if (kRes) {
      // the bounce needs the staging buffer: this warp's previous TMA store must have finished reading it
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      ADB_DBGE(3);
      slab_bounce<CS16>(qn, q, sbuf, lane);
      ADB_DBGE(13);
      // The next item's rows are requested only now, after this item's have left their registers: a request issued earlier
      // would share the scoreboard wait of the bounce above and expose its whole latency (measured: 2-4 k cycles per slab).
      sln += 2;
      while (sln >= P.n_slabs) { sln -= P.n_slabs; ++mtn; }
      load_residual<CS16, false, 0, CS16>(P, tc, rl, mtn, sln, !last, ew, lane, qn);            // first half of the loads
      ADB_DBGE(14);
    }
    tmem_ld_wait();
    if (kRes) { ADB_DBGE(15); }
    if (last) release_acc<kPair>(tempty, lane);
    slab_math<kAct, CS16, 0, CS16, kRes, kRes>(v, q, s_scale + ch0 + cl, s_shift + ch0 + cl, P.act, pk, sbuf, lane);
    if (kRes) load_residual<CS16, false, CS16, CS16 * 2>(P, tc, rl, mtn, sln, !last, ew, lane, qn);   // second half, after the arithmetic
    if (!kRes) {
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      ADB_DBGE(3);
      slab_stage<CS16>(pk, sbuf, lane);
    }
    ADB_DBGE(4);
    fence_proxy_async_smem();
    __syncwarp();
    ADB_DBGE(5);
    if (lane == 0) {   // the same thread owns this warp's bulk-group bookkeeping (commit / wait_group)
      tma_store_5d(tmOut, sbuf, st_c + cl, st_w, st_p, st_h + mt * P.TH, st_n);
      ADB_DBGE(16);
      tma_store_commit();
    }
    if (P.stat_out) {
      const int row = q_row0 + lane;
      const bool ok = (tc.h0 + mt * P.TH + (row >> P.tw_shift) < P.grid_h) && (tc.w0 + (row & (P.TW - 1)) < P.grid_w);
      const uint32_t valid_rows = __ballot_sync(0xffffffffu, ok);
      slab_stats<CS16>(P, sbuf, lane, valid_rows, (slot_base + (size_t)mt) * 4 + (size_t)ew, ch0 + cl);
    }
    ADB_DBGE(6);
    sl += 2;
    while (sl >= P.n_slabs) { sl -= P.n_slabs; ++mt; }
  }
  if (half >= items) release_acc<kPair>(tempty, lane);   // a warp without an item in this tile still owes its arrival
}

template <int kAct, int CS16, bool kPair>
__device__ __forceinline__ void feature_tile_any(const ConvK& P, const CUtensorMap* tmOut, const TileCoord& tc, uint32_t tfull,
                                                 uint32_t tfull_phase, uint32_t tempty, uint32_t tmem_tile, uint32_t sbuf, const float* s_scale,
                                                 const float* s_shift, int ew, int half, int lane, int& e_i, size_t slot_base) {
  if (P.residual) feature_tile<kAct, CS16, true, kPair>(P, tmOut, tc, tfull, tfull_phase, tempty, tmem_tile, sbuf, s_scale, s_shift, ew, half, lane, e_i, slot_base);
  else feature_tile<kAct, CS16, false, kPair>(P, tmOut, tc, tfull, tfull_phase, tempty, tmem_tile, sbuf, s_scale, s_shift, ew, half, lane, e_i, slot_base);
}

template <int kAct, bool kPair, int kEpi, bool kPre = false>
__global__ void __launch_bounds__(kPre ? kThreadsPre : kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ ConvK P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const SmemLayout L = smem_layout(P.a_slots, P.a_slot_bytes, P.b_slots, P.b_slot_bytes, P.slab_bytes, P.n_slab_bufs, P.cout_pad,
                                   kPre ? P.ctot : 0);
  const uint32_t a_base = base + L.a_off;
  const uint32_t b_base = base + L.b_off;
  const uint32_t slab_base = base + L.slab_off;
  float* s_scale = reinterpret_cast<float*>(base_ptr + L.scale_off);
  float* s_shift = s_scale + P.cout_pad;
  const uint32_t bar_base = base + L.bar_off;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (kMaxASlots + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (2 * kMaxASlots + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (2 * kMaxASlots + kMaxBSlots + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxASlots + 2 * kMaxBSlots + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxASlots + 2 * kMaxBSlots + 2 + a); };
  auto readyA = [&](int s) { return bar_base + 8u * (2 * kMaxASlots + 2 * kMaxBSlots + 6 + s); };   // (after the tmem pointer word)
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(base_ptr + L.bar_off + 8u * (2 * kMaxASlots + 2 * kMaxBSlots + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA-pair mode: the two CTAs of a cluster walk the same tile sequence; rank 0 (the leader) issues the MMAs.
  const int rank = kPair ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // the leader's barrier block as a shared::cluster address (TMA complete_tx / remote arrives of the peer go there)
  const uint32_t lead_bar_base = kPair ? mapa_shared(bar_base, 0) : bar_base;
  auto fullA_lead = [&](int s) { return lead_bar_base + 8u * s; };
  auto fullB_lead = [&](int s) { return lead_bar_base + 8u * (2 * kMaxASlots + s); };
  auto tempty_lead = [&](int a) { return lead_bar_base + 8u * (2 * kMaxASlots + 2 * kMaxBSlots + 2 + a); };

  // live image count of this launch (routed buckets carry it on the device)
  int n_eff = P.n;
  if (P.n_dev) n_eff = max(0, min(P.n, *P.n_dev - P.n_start));
  const int tiles_per_img = P.tiles_w * P.tiles_h * P.ngroups * P.n_tiles_n;
  const int total_tiles = n_eff * tiles_per_img;
  const int nchunks = P.chunks0 + P.chunks1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.a_slots; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < P.b_slots; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kPair ? 16 : 8); }   // 8 epilogue warps per CTA
    if (kPre) for (int s = 0; s < P.a_slots; ++s) mbar_init(readyA(s), 4);                                 // 4 transform warps
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) {
      tmem_alloc_pair(smem_u32((const void*)tmem_ptr_smem), (uint32_t)P.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)P.tmem_cols);
      tmem_relinquish();
    }
  }
  if (warp >= kEpiWarp0 && warp < kPreWarp0) {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < P.cout_pad; i += 256) {
      s_scale[i] = P.scale[i];
      s_shift[i] = P.shift[i];
    }
  }
  uint32_t* s_tap = reinterpret_cast<uint32_t*>(base_ptr + L.tap_off);   // A descriptor offset of every tap (16-byte units)
  if (warp == 1) {
    for (int i = lane; i < P.ngroups * kMaxTaps; i += 32) {
      const int g = i / kMaxTaps, tt = i - g * kMaxTaps;
      s_tap[i] = tt < P.ntaps ? ((uint32_t)P.taps[g][tt].shift_px * (uint32_t)P.row_bytes) >> 4 : 0u;
    }
  }
  float* s_pre = reinterpret_cast<float*>(base_ptr + L.pre_off);   // [ctot] scale then [ctot] shift
  if (kPre && warp >= kPreWarp0) {
    for (int i = threadIdx.x - kPreWarp0 * 32; i < P.ctot; i += 128) {
      s_pre[i] = P.pre_scale[i];
      s_pre[P.ctot + i] = P.pre_shift[i];
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // All three pipeline roles walk the same nest: tile -> channel chunk (src0 then src1) -> A load -> tap.
  if (warp == 0) {
    // ======================================================= A producer: one halo box per (chunk, A load)
    int slot = 0; uint32_t phase = 0; int dbg_i = 0;
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = decode_tile(P, t, rank);
      const int nal = P.n_aloads[tc.g];
      for (int c = 0; c < nchunks; ++c) {
        const bool s1 = c >= P.chunks0;
        const int coff = (s1 ? c - P.chunks0 : c) * P.Ck;
        const int pitch = s1 ? P.pitch1 : P.pitch0;
        const CUtensorMap* tm = s1 ? &tmA1 : &tmA0;
        for (int a = 0; a < nal; ++a) {
          const ALoad al = P.aloads[tc.g][a];
          mbar_wait(emptyA(slot), phase ^ 1u, P.err_flag, 1);
          ADB_DBG(0, dbg_i); ++dbg_i;
          if (elect_one()) {
            if (kPair) {   // both CTAs' halo boxes are counted on the leader's barrier
              if (leader) mbar_expect_tx(fullA(slot), 2u * (uint32_t)P.a_tx_bytes);
              tma_load_5d_pair(a_base + (uint32_t)slot * P.a_slot_bytes, tm, fullA_lead(slot), al.c_mul * pitch + coff,
                               tc.w0 + al.dw0, al.p, tc.h0 + al.dh0, tc.img);
            } else {
              mbar_expect_tx(fullA(slot), (uint32_t)P.a_tx_bytes);
              tma_load_5d(a_base + (uint32_t)slot * P.a_slot_bytes, tm, fullA(slot), al.c_mul * pitch + coff,
                          tc.w0 + al.dw0, al.p, tc.h0 + al.dh0, tc.img);
            }
          }
          __syncwarp();
          if (++slot == P.a_slots) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ======================================================= B producer: one weight box per (chunk, tap)
    int slot = 0; uint32_t phase = 0; int dbg_i = 0;
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = decode_tile(P, t, rank);
      const int brow = tc.g * P.cout_pad + tc.nt * P.BN + rank * (P.BN / P.ncta);   // pair: each CTA holds half of the N rows
      const int nal = P.n_aloads[tc.g];
      for (int c = 0; c < nchunks; ++c) {
        const int kc = (c >= P.chunks0 ? P.c0 + (c - P.chunks0) * P.Ck : c * P.Ck);   // channel offset in the concat
        for (int a = 0; a < nal; ++a) {
          const ALoad al = P.aloads[tc.g][a];
          for (int j0 = 0; j0 < al.tap_count; j0 += P.taps_per_slot) {
            const int nt = min(P.taps_per_slot, (int)al.tap_count - j0);
            mbar_wait(emptyB(slot), phase ^ 1u, P.err_flag, 5);
            ADB_DBG(1, dbg_i); ++dbg_i;
            if (elect_one()) {
              if (kPair) { if (leader) mbar_expect_tx(fullB(slot), 2u * (uint32_t)(nt * P.b_tx_bytes)); }
              else mbar_expect_tx(fullB(slot), (uint32_t)(nt * P.b_tx_bytes));
              for (int jj = 0; jj < nt; ++jj) {
                const TapK tk = P.taps[tc.g][al.tap_begin + j0 + jj];
                const uint32_t dst = b_base + (uint32_t)slot * P.b_slot_bytes + (uint32_t)jj * P.b_tap_stride;
                if (kPair) tma_load_2d_pair(dst, &tmB, fullB_lead(slot), tk.kidx * P.ctot + kc, brow);
                else tma_load_2d(dst, &tmB, fullB(slot), tk.kidx * P.ctot + kc, brow);
              }
            }
            __syncwarp();
            if (++slot == P.b_slots) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ======================================================= MMA issuer (whole warp waits, one elected lane issues; pair: leader CTA only)
    int sa = 0; uint32_t pa = 0;
    int sb = 0; uint32_t pb = 0;
    int acc = 0; uint32_t acc_phase = 0; int dbg_i = 0;
    const int ksteps_full = P.Ck / 16;
    const uint64_t desc_hi = make_kmajor_desc(0, P.row_bytes);     // (weights) everything but the start address
    // A views: same, except that the 8-row groups of an 8-pixel-wide 2-D tile are one halo row apart
    const uint64_t desc_hi_a = (desc_hi & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(((uint32_t)P.a_sbo_bytes >> 4) & 0x3FFF) << 32);
    IssueK K;
    K.a_hi = (uint32_t)(desc_hi_a >> 32); K.b_hi = (uint32_t)(desc_hi >> 32);
    K.sub_step = ((uint32_t)P.sub_px * (uint32_t)P.row_bytes) >> 4;
    K.b_tap_step = (uint32_t)P.b_tap_stride >> 4;
    K.idesc = P.idesc;
    const uint32_t lo_flags = (uint32_t)desc_hi;                  // LBO field
    const uint32_t a_lo0 = ((a_base & 0x3FFFFu) >> 4) | lo_flags, a_step = (uint32_t)P.a_slot_bytes >> 4;
    const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | lo_flags, b_step = (uint32_t)P.b_slot_bytes >> 4;
    uint32_t a_lo = a_lo0, b_lo = b_lo0;
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = decode_tile(P, t, 0);
      const int nal = P.n_aloads[tc.g];
      const uint32_t* taps_g = s_tap + tc.g * kMaxTaps;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u, P.err_flag, 2);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(acc * P.MT * P.bn_cols);
      const uint32_t d1 = d_base + (uint32_t)P.bn_cols;
      uint32_t accumulate = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int ksteps = c == P.chunks0 - 1 ? P.ks_last0 : (c == nchunks - 1 ? P.ks_last1 : ksteps_full);
        for (int a = 0; a < nal; ++a) {
          const int tap_begin = P.aloads[tc.g][a].tap_begin, tap_count = P.aloads[tc.g][a].tap_count;
          mbar_wait(kPre ? readyA(sa) : fullA(sa), pa, P.err_flag, 3);
          for (int j0 = 0; j0 < tap_count; j0 += P.taps_per_slot) {
            const int nt = min(P.taps_per_slot, tap_count - j0);
            const bool last_of_a = j0 + nt >= tap_count;
            mbar_wait(fullB(sb), pb, P.err_flag, 6);
            tc_fence_after();
            ADB_DBG(2, dbg_i);
            if (elect_one()) {   // elect.sync lets the compiler keep the UTCHMMA stream in straight-line uniform code
              const uint32_t* tp = taps_g + tap_begin + j0;
              if (P.MT == 2) {
                if (ksteps == 4) issue_taps<kPair, 2, 4>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
                else if (ksteps == 3) issue_taps<kPair, 2, 3>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
                else if (ksteps == 2) issue_taps<kPair, 2, 2>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
                else issue_taps<kPair, 2, 1>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
              } else {
                if (ksteps == 4) issue_taps<kPair, 1, 4>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
                else if (ksteps == 3) issue_taps<kPair, 1, 3>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
                else if (ksteps == 2) issue_taps<kPair, 1, 2>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
                else issue_taps<kPair, 1, 1>(K, d_base, d1, a_lo, b_lo, tp, nt, accumulate);
              }
              const bool tile_done = c == nchunks - 1 && a == nal - 1 && last_of_a;
              if (kPair) {   // multicast commits: the slot / accumulator barriers of BOTH CTAs
                umma_commit_pair(emptyB(sb));
                if (last_of_a) umma_commit_pair(emptyA(sa));
                if (tile_done) umma_commit_pair(tfull_bar(acc));
              } else {
                umma_commit(emptyB(sb));                                   // weight slot free once these MMAs have read it
                if (last_of_a) umma_commit(emptyA(sa));                    // halo slot free after its last tap
                if (tile_done) umma_commit(tfull_bar(acc));
              }
            }
            __syncwarp();
            ADB_DBG(3, dbg_i); ++dbg_i;
            accumulate = 1;
            b_lo += b_step;
            if (++sb == P.b_slots) { sb = 0; pb ^= 1u; b_lo = b_lo0; }
          }
          a_lo += a_step;
          if (++sa == P.a_slots) { sa = 0; pa ^= 1u; a_lo = a_lo0; }
        }
      }
      if (++acc == P.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (kPre && warp >= kPreWarp0) {
    // ======================================================= operand transform (4 warps): the consumer's pre-activation
    // relu(x*scale[c] + shift[c]) (DenseNet norm1/relu1 ahead of conv1 — a different affine of the same concat for every
    // layer) is applied to each A box in shared memory between the TMA fill and the MMAs, so the normalised copy of the
    // concat never goes through HBM.  16-byte chunks; the channel of a chunk follows from the TMA swizzle.
    int slot = 0; uint32_t phase = 0;
    const int tid = threadIdx.x - kPreWarp0 * 32;
    const uint32_t swz_mask = ((uint32_t)P.row_bytes >> 4) - 1u;
    const float* pre_sc = s_pre;
    const float* pre_sh = s_pre + P.ctot;
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = decode_tile(P, t, rank);
      const int nal = P.n_aloads[tc.g];
      for (int c = 0; c < nchunks; ++c) {
        const int kc = (c >= P.chunks0 ? P.c0 + (c - P.chunks0) * P.Ck : c * P.Ck);
        for (int a = 0; a < nal; ++a) {
          mbar_wait(fullA(slot), phase, P.err_flag, 7);
          const uint32_t sbase = a_base + (uint32_t)slot * P.a_slot_bytes;
          // a thread's chunks are 2048 B apart (16 or 32 whole rows), so their swizzle phase — hence their channel group —
          // is the same for every chunk it touches: the affine is loaded once per box
          const uint32_t o0 = (uint32_t)tid * 16u;
          const uint32_t lo = o0 ^ (((o0 >> 7) & swz_mask) << 4);
          const int ch = min(kc + (int)((lo & ((uint32_t)P.row_bytes - 1u)) >> 1), P.ctot - 8);   // (ragged tail: never read by an MMA)
          const float4 s0 = *reinterpret_cast<const float4*>(pre_sc + ch), s1 = *reinterpret_cast<const float4*>(pre_sc + ch + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(pre_sh + ch), h1 = *reinterpret_cast<const float4*>(pre_sh + ch + 4);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          const uint32_t nbytes = (uint32_t)P.a_tx_bytes;
          for (uint32_t o = o0; o < nbytes; o += 4u * 2048u) {        // four independent chunks in flight per thread
            uint32_t v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t a = o + (uint32_t)u * 2048u;
              if (a < nbytes)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]) : "r"(sbase + a) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t a = o + (uint32_t)u * 2048u;
              if (a < nbytes) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&v[u][i]);
                  v[u][i] = pack_bf16x2_relu(fmaf(__low2float(b2), sc[2 * i], sh[2 * i]), fmaf(__high2float(b2), sc[2 * i + 1], sh[2 * i + 1]));
                }
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sbase + a), "r"(v[u][0]), "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3]) : "memory");
              }
            }
          }
          fence_proxy_async_smem();       // generic-proxy writes -> visible to the tensor core's async-proxy operand reads
          __syncwarp();
          if (lane == 0) mbar_arrive(readyA(slot));
          if (++slot == P.a_slots) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kPreWarp0) {
    // ======================================================= epilogue (8 warps; thread = pixel row of the tile)
    // Warps 4-7 and 8-11 both cover the four TMEM lane quarters (a warp may only read lanes 32*(warp%4)..+31); the two
    // warps of a quarter split the tile's work items, which doubles the instruction throughput of an otherwise
    // latency-bound single-warp-per-scheduler epilogue.
    const int ewi = warp - kEpiWarp0;                   // 0..7
    const int ew = ewi & 3;                             // TMEM lane quarter
    const int half = ewi >> 2;                          // which alternate work items of a tile this warp takes
    const int et = ew * 32 + lane;                      // 0..127 == TMEM lane == tile row
    const int th_l = et >> P.tw_shift, tw_l = et & (P.TW - 1);
    int acc = 0; uint32_t acc_phase = 0; int dbg_i = 0; int e_i = 0;
    const uint32_t sbuf = slab_base + (uint32_t)ewi * (uint32_t)(P.slab_bytes >> 2);   // this warp's staging buffer
    float dotw[32];                                     // DOT: the 1x1 head's weights over the (<= 32) conv channels
    if (kEpi == ADB_EPI_DOT) {
#pragma unroll
      for (int i = 0; i < 32; ++i) dotw[i] = i < P.BN ? __ldg(P.dot_w + i) : 0.f;
    }
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = decode_tile(P, t, rank);
      ADB_DBGE(12);
      const uint32_t tmem_tile = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * P.MT * P.bn_cols);
      if (kEpi == ADB_EPI_FEATURE) {
        // (kPre kernels run 512 threads = 128 registers each: their epilogue uses the 32-channel slab at most)
        // partial-statistics slot of this CTA's first sub-tile: one slot per (pixel tile, CTA of a pair, sub-tile, lane quarter);
        // the N tiles of one pixel tile share it (they own disjoint channel ranges)
        const size_t slot_base = (((size_t)t / (size_t)P.n_tiles_n) * (size_t)P.ncta + (size_t)rank) * (size_t)P.MT;
        if (!kPre && P.Cs == 64) feature_tile_any<kAct, 4, kPair>(P, &tmOut, tc, tfull_bar(acc), acc_phase, kPair ? tempty_lead(acc) : tempty_bar(acc), tmem_tile, sbuf, s_scale, s_shift, ew, half, lane, e_i, slot_base);
        else if (P.Cs == 32) feature_tile_any<kAct, 2, kPair>(P, &tmOut, tc, tfull_bar(acc), acc_phase, kPair ? tempty_lead(acc) : tempty_bar(acc), tmem_tile, sbuf, s_scale, s_shift, ew, half, lane, e_i, slot_base);
        else feature_tile_any<kAct, 1, kPair>(P, &tmOut, tc, tfull_bar(acc), acc_phase, kPair ? tempty_lead(acc) : tempty_bar(acc), tmem_tile, sbuf, s_scale, s_shift, ew, half, lane, e_i, slot_base);
        if (ewi == 0) { ADB_DBG(4, dbg_i); }
      } else {
        // DOT / IMAGE: one work item per sub-tile; `half` takes sub-tile `half`
        const int mt = half;
        const int h = tc.h0 + mt * P.TH + th_l;
        const int w = tc.w0 + tw_l;
        const bool inb = mt < P.MT && (h < P.grid_h) && (w < P.grid_w);
        // IMAGE: the hazy pixels / guidance this thread combines with its accumulator row do not depend on the MMAs: fetch
        // them before waiting for the accumulator so their latency hides under this tile's main loop.
        float xin[3] = {0.f, 0.f, 0.f};
        float gd = 1.f, alpha = 0.f;
        size_t img_o = 0;
        const size_t plane = (size_t)P.grid_h * P.grid_w;
        if (kEpi == ADB_EPI_IMAGE && inb) {
          const int pos = P.n_start + tc.img;
          const size_t row = P.img_index ? (size_t)__ldg(P.img_index + pos) : (size_t)pos;
          if (P.img_mode == ADB_IMG_BLEND) alpha = __ldg(P.img_alpha);
          img_o = row * 3 * plane + (size_t)h * P.grid_w + w;
          if (P.img_mode == ADB_IMG_GUIDED) gd = __ldg(P.img_guidance + ((size_t)tc.img * P.grid_h + h) * P.grid_w + w);
#pragma unroll
          for (int c = 0; c < 3; ++c) xin[c] = __ldg(P.img_x + img_o + c * plane);
        }
        mbar_wait(tfull_bar(acc), acc_phase, P.err_flag, 4);
        tc_fence_after();
        if (ewi == 0) { ADB_DBG(4, dbg_i); }
        if (mt < P.MT) {
          float v[16];
          tmem_ld16(tmem_tile + (uint32_t)(mt * P.bn_cols), v);
          tmem_ld_wait();
          if (kEpi == ADB_EPI_DOT) {
            float g = P.dot_b;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float y = apply_act(fmaf(v[i], s_scale[i], s_shift[i]), P.act);
              g = fmaf(y, dotw[i], g);
            }
            if (P.BN > 16) {                            // second 16-column group (24- / 32-channel heads)
              tmem_ld16(tmem_tile + (uint32_t)(mt * P.bn_cols + 16), v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float y = apply_act(fmaf(v[i], s_scale[16 + i], s_shift[16 + i]), P.act);
                g = fmaf(y, dotw[16 + i], g);
              }
            }
            g = 1.f / (1.f + __expf(-g));
            if (inb) P.dot_out[((size_t)tc.img * P.grid_h + h) * P.grid_w + w] = g;
          } else if (inb) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float y = apply_act(fmaf(v[c], s_scale[c], s_shift[c]), P.act);
              float out;
              if (P.img_mode == ADB_IMG_BLEND) out = (1.f - alpha) * xin[c] + alpha * y;
              else out = fminf(fmaxf(xin[c] + y * gd, 0.f), 1.f);
              P.img_out[img_o + c * plane] = out;
            }
          }
        }
      }
      // accumulator stage fully read by this warp -> hand it back to the MMA issuer (8 arrivals per CTA complete the phase)
      if (ewi == 0) { ADB_DBG(5, dbg_i); ++dbg_i; }
      ADB_DBGE(9);
      if (kEpi != ADB_EPI_FEATURE) release_acc<kPair>(kPair ? tempty_lead(acc) : tempty_bar(acc), lane);   // (FEATURE: inside feature_tile)
      ADB_DBGE(11);
      if (++acc == P.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ---------------------------------------------------------- teardown
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // pair: neither CTA may retire while its peer can still signal it
  if (warp == 2) { if (kPair) tmem_dealloc_pair(tmem_base, (uint32_t)P.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
}

struct RawTap { int c_mul, p, dw, dh, kidx; };

int build(const adb_conv_desc* d, ConvK& P, int& out_h, int& out_w, int& ktot, int& box_w, int& box_h) {
  memset(&P, 0, sizeof(P));
  ADB_REQUIRE(d != nullptr, "adb_conv2d: null descriptor");
  ADB_REQUIRE(d->src0 && d->c0 > 0 && d->c0_pitch >= d->c0 && d->c0_pitch % 8 == 0, "adb_conv2d: bad src0 (c0=%d pitch=%d)", d->c0, d->c0_pitch);
  ADB_REQUIRE((d->src1 == nullptr) == (d->c1 == 0), "adb_conv2d: src1/c1 mismatch");
  if (d->src1) ADB_REQUIRE(d->c1 > 0 && d->c1_pitch >= d->c1 && d->c1_pitch % 8 == 0, "adb_conv2d: bad src1 (c1=%d pitch=%d)", d->c1, d->c1_pitch);
  ADB_REQUIRE(d->n > 0 && d->h_in > 0 && d->w_in > 0, "adb_conv2d: bad n/h/w %d/%d/%d", d->n, d->h_in, d->w_in);
  ADB_REQUIRE(d->w_packed && d->scale && d->shift, "adb_conv2d: null weights/scale/shift");
  ADB_REQUIRE(d->cout > 0 && d->cout_pad >= d->cout && d->cout_pad % 16 == 0 && d->cout_pad <= 512, "adb_conv2d: bad cout %d / cout_pad %d", d->cout, d->cout_pad);

  int Ck = pick_chunk(d->c0);
  if (d->src1) Ck = std::min(Ck, pick_chunk(d->c1));
  ADB_REQUIRE(Ck > 0 && d->c0 % Ck == 0 && (d->c1 % Ck) == 0, "adb_conv2d: channel counts %d/%d must be multiples of 16", d->c0, d->c1);
  // Ragged chunking: channel counts such as 48 / 96 / 160 would force 16- or 32-channel boxes (32/64-byte TMA rows and twice
  // to four times the barrier round trips per MMA).  Use 64-channel boxes anyway and issue only the K steps that hold real
  // channels in each source's last chunk; whatever the box fetches beyond them (zero fill, or a neighbour's channels) is
  // never read by an MMA.  Not for the space-to-depth view, whose channel axis interleaves the two column phases.
  // 32-channel sources keep their 32-channel boxes: a half-empty 64-channel box was tried (to give the Light branch's A
  // operand 128-byte rows) and measured slower — 0.319 vs 0.279 ms on light_32_3x3, the doubled halo slot also no longer
  // fits shared memory at 1024x2048 (profiles/r1m/prof_conv_ragged32.txt).
  const bool ragged = Ck < 64 && d->kind != ADB_CONV_S2 && !(d->tune_flags & 128) &&
                      (d->c0 >= 48 || d->c1 >= 48);
  if (ragged) Ck = 64;
  P.Ck = Ck; P.row_bytes = Ck * 2;
  P.chunks0 = (d->c0 + Ck - 1) / Ck; P.chunks1 = (d->c1 + Ck - 1) / Ck;
  P.ks_last0 = (d->c0 - (P.chunks0 - 1) * Ck) / 16;
  P.ks_last1 = d->c1 ? (d->c1 - (P.chunks1 - 1) * Ck) / 16 : 0;
  P.pitch0 = d->c0_pitch; P.pitch1 = d->src1 ? d->c1_pitch : d->c0_pitch;
  P.ctot = d->c0 + d->c1;
  P.c0 = d->c0;
  const bool pre = d->pre_scale != nullptr;
  if (pre) {
    ADB_REQUIRE(d->pre_shift != nullptr, "adb_conv2d: pre_scale and pre_shift go together");
    ADB_REQUIRE(d->kind == ADB_CONV_S1 && d->kh == 1 && d->kw == 1 && d->epi == ADB_EPI_FEATURE,
                "adb_conv2d: the fused pre-activation is built for 1x1 stride-1 FEATURE convs (zero padding of a k x k conv applies "
                "to the ACTIVATED input, which an in-operand transform cannot reproduce)");
    ADB_REQUIRE(P.ctot <= 4096, "adb_conv2d: pre-activation tables hold at most 4096 channels");
  }
  P.pre_scale = d->pre_scale; P.pre_shift = d->pre_shift;

  // ---- raw taps per group: (space-to-depth phases, pixel offsets, position in the packed K order)
  RawTap raw[kMaxGroups][kMaxTaps];
  if (d->kind == ADB_CONV_S1) {
    ADB_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= kMaxTaps, "adb_conv2d: %dx%d taps unsupported (max %d)", d->kh, d->kw, kMaxTaps);
    // 'same' convolution: the padding is implied by the (odd) kernel extents, so a kh x 1 stem conv pads rows only
    ADB_REQUIRE(d->kh % 2 == 1 && d->kw % 2 == 1, "adb_conv2d: stride-1 convs need odd kernel extents (got %dx%d)", d->kh, d->kw);
    const int pad_h = (d->kh - 1) / 2, pad_w = (d->kw - 1) / 2;
    ADB_REQUIRE(d->pad == std::max(pad_h, pad_w), "adb_conv2d: stride-1 convs must be 'same' (pad=(k-1)/2, got %d)", d->pad);
    out_h = d->h_in; out_w = d->w_in;
    P.ngroups = 1; P.ntaps = d->kh * d->kw;
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) raw[0][r * d->kw + s] = {0, 0, s - pad_w, r - pad_h, r * d->kw + s};
    P.grid_h = out_h; P.grid_w = out_w;
  } else if (d->kind == ADB_CONV_S2) {
    ADB_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= kMaxTaps, "adb_conv2d: %dx%d taps unsupported", d->kh, d->kw);
    ADB_REQUIRE(d->h_in % 2 == 0 && d->w_in % 2 == 0, "adb_conv2d: stride-2 convs need even H/W (got %dx%d)", d->h_in, d->w_in);
    out_h = (d->h_in + 2 * d->pad - d->kh) / 2 + 1;
    out_w = (d->w_in + 2 * d->pad - d->kw) / 2 + 1;
    ADB_REQUIRE(out_h == d->h_in / 2 && out_w == d->w_in / 2, "adb_conv2d: stride-2 conv must halve H/W");
    P.ngroups = 1; P.ntaps = d->kh * d->kw;
    auto fl2 = [](int u) { return (u >= 0) ? u / 2 : -((-u + 1) / 2); };
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) {
        const int u = r - d->pad, v = s - d->pad;
        raw[0][r * d->kw + s] = {v - 2 * fl2(v), u - 2 * fl2(u), fl2(v), fl2(u), r * d->kw + s};
      }
    P.grid_h = out_h; P.grid_w = out_w;
  } else if (d->kind == ADB_CONV_K4_S2D) {
    ADB_REQUIRE(d->kh == 4 && d->kw == 4, "adb_conv2d: the space-to-depth stem form is a 4x4 tap grid");
    out_h = d->h_in; out_w = d->w_in;
    P.ngroups = 1; P.ntaps = 16;
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 4; ++s) raw[0][r * 4 + s] = {0, 0, s - 2, r - 2, r * 4 + s};
    P.grid_h = out_h; P.grid_w = out_w;
  } else if (d->kind == ADB_CONVT_4X4S2) {
    out_h = d->h_in * 2; out_w = d->w_in * 2;
    P.ngroups = 4; P.ntaps = 4;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) raw[a * 2 + b][i * 2 + j] = {0, 0, b ? 1 - j : -j, a ? 1 - i : -i, i * 2 + j};
    P.grid_h = d->h_in; P.grid_w = d->w_in;
  } else {
    return adbh::fail(ADB_ERR_INVALID, "adb_conv2d: unknown kind %d", d->kind);
  }
  ktot = P.ntaps * P.ctot;

  // ---- N tiling
  P.BN = d->cout_pad;
  P.n_tiles_n = 1;
  if (d->cout_pad > 256) {
    ADB_REQUIRE(d->cout_pad % 32 == 0, "adb_conv2d: cout_pad %d > 256 must split evenly", d->cout_pad);
    P.n_tiles_n = 2; P.BN = d->cout_pad / 2;
  }
  ADB_REQUIRE(P.BN % 16 == 0 && P.BN >= 16 && P.BN <= 256, "adb_conv2d: N tile %d invalid", P.BN);
  P.cout_pad = d->cout_pad;
  P.bn_cols = (2 * round_up(P.BN, 32) <= 512) ? round_up(P.BN, 32) : P.BN;

  // ---- pixel tile
  int TW = 128;
  while (TW > P.grid_w && TW > 8) TW >>= 1;
  // 2-D tiles (16 rows x 8 pixels; tune_flags bit 8): all taps share one halo box that over-fetches 1.4x (1.33x at MT = 2)
  // instead of the one-row tile's 2x; each 8-row group of the UMMA operand is one image row of the tile, one halo row
  // (TW + 2 pixels) from the next, which the descriptor's stride-byte-offset expresses directly.  Measured (tools/ab_conv.py,
  // same process, interleaved): no gain on the operand-light shapes it was meant for (Light 32->32, Medium 64->64, DenseNet
  // 128->32 are epilogue-latency bound, not fetch bound) and 6-8 % slower with a residual (less coalesced residual rows),
  // so it is opt-in only.
  bool tile2d = (d->tune_flags & 256) && d->kind == ADB_CONV_S1 && d->kw > 1 && P.grid_w >= 8;
  if (tile2d) TW = 8;
  P.TW = TW; P.TH = 128 / TW;
  long long sub_tiles = (long long)d->n * ((P.grid_h + P.TH - 1) / P.TH) * ((P.grid_w + TW - 1) / TW) * P.n_tiles_n * P.ngroups;
  // CTA-pair mode (cta_group::2): halves the weight-operand traffic per CTA (operand reads and TMA fill), which is what
  // bounds the 1-CTA kernel.  Worth it once the weights are a real share of the shared-memory traffic
  // (N tile >= 48, or >= 32 with a deep K; more than one tap) and there are enough tiles to fill 74 pairs for a few waves.
  // tune_flags bit 4 forces it on, bit 5 forces it off.
  int ncta = ((P.BN >= 48 || (P.BN >= 32 && P.ctot >= 128)) && P.ntaps > 1 && sub_tiles >= 8LL * 148) ? 2 : 1;
  if (d->tune_flags & 16) ncta = 2;
  if (d->tune_flags & 32) ncta = 1;
  if (pre) ncta = 1;
  P.ncta = ncta;
  // 256 output channels of a CTA pair run as two 128-channel N tiles: two sub-tiles per CTA then fit TMEM double-buffered
  // (2 x 2 x 128 columns), which beats one 256-column tile per stage by 11 % (0.844 -> 0.754 ms on med_256_3x3,
  // profiles/r2/prof_nsplit.txt).  The same split loses at 192 and 128 (operand A is fetched twice for too little gain) and
  // without the pair (1.06 -> 1.21 ms).  tune_flags bit 10 forces a split, bit 11 forbids it.
  const bool split_n = (d->tune_flags & 1024) || (P.BN == 256 && ncta == 2 && !(d->tune_flags & 2048));
  if (split_n && P.BN % 32 == 0 && P.BN >= 64) {
    P.n_tiles_n *= 2; P.BN /= 2; sub_tiles *= 2;
    P.bn_cols = (2 * round_up(P.BN, 32) <= 512) ? round_up(P.BN, 32) : P.BN;
  }
  // sub-tiles per CTA: two share every weight box when their accumulators fit TMEM; measured (profiles/r1_pair_sweep.txt):
  // pairs prefer MT = 2 even single-buffered (N = 192), except at N = 256 where double buffering wins.
  int mt = d->tune_mt > 0 ? d->tune_mt : (ncta == 2 ? (P.bn_cols <= 192 ? 2 : 1) : (P.bn_cols <= 128 ? 2 : 1));
  ADB_REQUIRE(mt == 1 || mt == 2, "adb_conv2d: tune_mt must be 1 or 2");
  if (mt * P.bn_cols > 512) mt = 1;
  if (sub_tiles < 2LL * 148 * mt * ncta) mt = 1;  // keep the SMs busy
  P.MT = mt;
  P.tiles_w = (P.grid_w + TW - 1) / TW;
  P.tiles_h = (P.grid_h + P.TH * mt * ncta - 1) / (P.TH * mt * ncta);
  P.fd_nt = make_fastdiv(P.n_tiles_n); P.fd_g = make_fastdiv(P.ngroups);
  P.fd_tw = make_fastdiv(P.tiles_w); P.fd_th = make_fastdiv(P.tiles_h);
  {
    const unsigned long long all_tiles = (unsigned long long)d->n * P.tiles_w * P.tiles_h * P.ngroups * P.n_tiles_n;
    const unsigned long long dmax = (unsigned long long)std::max(std::max(P.tiles_w, P.tiles_h), 4);
    ADB_REQUIRE(all_tiles * dmax < (1ULL << 32), "adb_conv2d: %llu tiles exceed the tile-decode range; split the batch", all_tiles);
  }
  P.tw_shift = 0;
  while ((1 << P.tw_shift) < TW) ++P.tw_shift;
  int acc = std::min(512 / (mt * P.bn_cols), 2);
  if (d->tune_acc_stages > 0) acc = std::min(acc, d->tune_acc_stages);
  ADB_REQUIRE(acc >= 1, "adb_conv2d: accumulators do not fit TMEM");
  P.acc_stages = acc;
  P.tmem_cols = pow2_at_least(acc * mt * P.bn_cols);
  P.idesc = make_idesc_bf16(128u * (uint32_t)P.ncta, (uint32_t)P.BN);
  P.desc_base_offset = (d->tune_flags & 2) ? 1 : 0;

  // ---- A loads: taps of one (c_mul, p) phase share a halo box when the 128-row views stay contiguous in it
  //      (tile = one image row, or no column shifts); otherwise each tap is its own box.
  int eh = 0, ew = 0;
  for (int g = 0; g < P.ngroups; ++g) {
    int nal = 0, ntp = 0;
    bool used[kMaxTaps] = {false};
    for (int t0 = 0; t0 < P.ntaps; ++t0) {
      if (used[t0]) continue;
      const RawTap& f = raw[g][t0];
      int dw_min = f.dw, dw_max = f.dw, dh_min = f.dh, dh_max = f.dh;
      int members[kMaxTaps], nm = 0;
      for (int t1 = t0; t1 < P.ntaps; ++t1) {
        const RawTap& q = raw[g][t1];
        if (used[t1] || q.c_mul != f.c_mul || q.p != f.p) continue;
        members[nm++] = t1;
        dw_min = std::min(dw_min, q.dw); dw_max = std::max(dw_max, q.dw);
        dh_min = std::min(dh_min, q.dh); dh_max = std::max(dh_max, q.dh);
      }
      const bool can_share = !(d->tune_flags & 1) && (P.TH == 1 || dw_max == dw_min || tile2d);
      if (!can_share) { nm = 1; members[0] = t0; dw_min = dw_max = f.dw; dh_min = dh_max = f.dh; }
      ADB_REQUIRE(nal < kMaxALoads, "adb_conv2d: too many A loads");
      ALoad& al = P.aloads[g][nal++];
      al.c_mul = (int16_t)f.c_mul; al.p = (int8_t)f.p; al.dw0 = (int8_t)dw_min; al.dh0 = (int8_t)dh_min;
      al.tap_begin = (uint8_t)ntp; al.tap_count = (uint8_t)nm;
      eh = std::max(eh, dh_max - dh_min); ew = std::max(ew, dw_max - dw_min);
      for (int k = 0; k < nm; ++k) {
        used[members[k]] = true;
        // shift is finalised below once the halo width is known; stash (ddh, ddw) for now
        P.taps[g][ntp].shift_px = (uint16_t)(((raw[g][members[k]].dh - dh_min) << 8) | (raw[g][members[k]].dw - dw_min));
        P.taps[g][ntp].kidx = (uint16_t)raw[g][members[k]].kidx;
        ++ntp;
      }
    }
    P.n_aloads[g] = nal;
  }
  P.halo_w = P.TW + ew;
  P.sub_px = P.TH * P.halo_w;
  P.a_sbo_bytes = (tile2d && ew > 0) ? P.halo_w * P.row_bytes : 8 * P.row_bytes;
  for (int g = 0; g < P.ngroups; ++g)
    for (int t = 0; t < P.ntaps; ++t) {
      const int ddh = P.taps[g][t].shift_px >> 8, ddw = P.taps[g][t].shift_px & 0xFF;
      P.taps[g][t].shift_px = (uint16_t)(ddh * P.halo_w + ddw);
    }
  box_w = P.halo_w;
  box_h = P.TH * mt + eh;
  ADB_REQUIRE(box_w <= 256 && box_h <= 256, "adb_conv2d: halo box %dx%d exceeds the TMA box limit", box_w, box_h);

  // ---- smem
  P.a_tx_bytes = box_w * box_h * P.row_bytes;
  P.b_tx_bytes = (P.BN / P.ncta) * P.row_bytes;     // pair: each CTA loads (and holds) half of the N rows
  P.a_slot_bytes = round_up(P.a_tx_bytes, 1024);
  P.b_tap_stride = round_up(P.b_tx_bytes, 1024);
  {
    // small weight boxes are batched: one mbarrier round-trip then covers several taps' worth of MMAs
    int max_taps = 1;
    for (int g = 0; g < P.ngroups; ++g)
      for (int a = 0; a < P.n_aloads[g]; ++a) max_taps = std::max(max_taps, (int)P.aloads[g][a].tap_count);
    int tb = std::max(1, (16 * 1024) / P.b_tap_stride);
    if (d->tune_flags & 8) tb = 1;
    P.taps_per_slot = std::min(tb, max_taps);
  }
  P.b_slot_bytes = P.taps_per_slot * P.b_tap_stride;
  if (d->epi == ADB_EPI_FEATURE) {
    P.Cs = pick_chunk(P.BN);
    if (pre) P.Cs = std::min(P.Cs, 32);
    P.n_slabs = P.BN / P.Cs;
    P.slab_bytes = 128 * P.Cs * 2;
  } else {
    ADB_REQUIRE((P.BN == 16 || (d->epi == ADB_EPI_DOT && P.BN == 32)) && P.n_tiles_n == 1,
                "adb_conv2d: the IMAGE epilogue needs cout_pad == 16, the DOT epilogue cout_pad 16 or 32 (got %d)", P.BN);
    P.Cs = 16; P.n_slabs = 0; P.slab_bytes = 1024;
  }
  adbh::DeviceInfo di;
  int st = adbh::device_info(&di);
  if (st != ADB_OK) return st;
  const int budget = di.max_smem_optin - 1024;  // alignment slack
  // Output staging: one buffer of 32 rows x Cs channels per epilogue warp (8 warps = 2 slabs).  Weights ring first
  // (>= 3 boxes), then up to 4 halo slots, then the rest back to the weights ring; the slab narrows from 64 to 32 channels
  // when that buys the third weight box.
  P.n_slab_bufs = 2;
  int a_slots = 2, b_slots = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const SmemLayout fixed = smem_layout(0, 0, 0, 0, P.slab_bytes, P.n_slab_bufs, P.cout_pad, pre ? P.ctot : 0);
    const int avail = budget - (int)fixed.total;
    a_slots = 2;
    b_slots = (avail - a_slots * P.a_slot_bytes) / P.b_slot_bytes;
    const bool can_narrow = pass == 0 && d->epi == ADB_EPI_FEATURE && P.Cs == 64;
    if (b_slots >= 3 || !can_narrow) {
      ADB_REQUIRE(b_slots >= 2, "adb_conv2d: pipeline does not fit shared memory (A %d B, B %d B)", P.a_slot_bytes, P.b_slot_bytes);
      while (a_slots < 4 && (a_slots + 1) * P.a_slot_bytes + 4 * P.b_slot_bytes <= avail) ++a_slots;
      b_slots = std::min(kMaxBSlots, (avail - a_slots * P.a_slot_bytes) / P.b_slot_bytes);
      break;
    }
    P.Cs = 32; P.n_slabs = P.BN / 32; P.slab_bytes = 128 * 32 * 2;
  }
  if (d->tune_stages > 0) { a_slots = std::min(a_slots, std::max(2, d->tune_stages)); b_slots = std::min(b_slots, std::max(2, d->tune_stages)); }
  P.a_slots = a_slots; P.b_slots = b_slots;

  // ---- batch / epilogue
  P.n = d->n; P.n_start = d->n_start; P.n_dev = d->n_dev;
  P.epi = d->epi; P.act = d->act;
  P.scale = d->scale; P.shift = d->shift;
  if (d->epi == ADB_EPI_FEATURE) {
    ADB_REQUIRE(d->dst && d->dst_pitch % 8 == 0 && d->dst_c_off >= 0 && d->dst_c_off + d->cout_pad <= d->dst_pitch,
                "adb_conv2d: dst pitch %d cannot hold channels [%d, %d)", d->dst_pitch, d->dst_c_off, d->dst_c_off + d->cout_pad);
    ADB_REQUIRE(d->dst_c_off % 8 == 0, "adb_conv2d: dst_c_off must be a multiple of 8");
    if (d->residual) {
      ADB_REQUIRE(d->kind != ADB_CONVT_4X4S2, "adb_conv2d: residual unsupported for ConvTranspose");
      ADB_REQUIRE(d->res_pitch >= d->cout_pad && d->res_pitch % 8 == 0, "adb_conv2d: residual pitch %d too small", d->res_pitch);
    }
    P.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
    P.res_pitch = d->res_pitch;
    if (d->stat_out) {
      ADB_REQUIRE(d->stat_mode >= 0 && d->stat_mode <= 2, "adb_conv2d: stat_mode must be 1 (sum, sum of squares) or 2 (sum, max)");
      ADB_REQUIRE(d->kind != ADB_CONVT_4X4S2 && !pre, "adb_conv2d: stat_out is built for plain / stride-2 FEATURE convs");
      P.stat_out = d->stat_out; P.stat_mode = d->stat_mode == 2 ? 2 : 1;
    }
    for (int g = 0; g < P.ngroups; ++g) {
      if (d->kind == ADB_CONVT_4X4S2) { P.out_c_off[g] = (g & 1) * d->dst_pitch + d->dst_c_off; P.out_p[g] = g >> 1; }
      else { P.out_c_off[g] = d->dst_c_off; P.out_p[g] = 0; }
    }
  } else if (d->epi == ADB_EPI_DOT) {
    ADB_REQUIRE(!d->stat_out, "adb_conv2d: stat_out needs the FEATURE epilogue");
    ADB_REQUIRE(d->dot_w && d->dot_out && d->kind != ADB_CONVT_4X4S2, "adb_conv2d: DOT epilogue needs dot_w/dot_out");
    P.dot_w = d->dot_w; P.dot_b = d->dot_b; P.dot_out = d->dot_out;
  } else if (d->epi == ADB_EPI_IMAGE) {
    ADB_REQUIRE(d->img_x && d->img_out && d->cout == 3 && d->kind == ADB_CONV_S1, "adb_conv2d: IMAGE epilogue needs img_x/img_out, cout == 3");
    ADB_REQUIRE(d->img_mode != ADB_IMG_GUIDED || d->img_guidance, "adb_conv2d: GUIDED needs img_guidance");
    ADB_REQUIRE(d->img_mode != ADB_IMG_BLEND || d->img_alpha, "adb_conv2d: BLEND needs img_alpha");
    P.img_mode = d->img_mode; P.img_x = d->img_x; P.img_out = d->img_out; P.img_index = d->img_index;
    P.img_guidance = d->img_guidance; P.img_alpha = d->img_alpha;
  } else {
    return adbh::fail(ADB_ERR_INVALID, "adb_conv2d: unknown epilogue %d", d->epi);
  }
  return ADB_OK;
}

}  // namespace

namespace adbc {
bool roll_eligible(const adb_conv_desc* d);                 // conv_roll.cu
int conv_roll_launch(const adb_conv_desc* d, void* stream);
}

static long long* g_dbg = nullptr;

namespace adbc {
// the (lazily allocated, zeroed on `stream`) debug buffer adb_debug_timeline() reads back
int debug_buffer(long long** out, void* stream) {
  if (!g_dbg) ADB_CUDA_OK(cudaMalloc(&g_dbg, 6 * 256 * sizeof(long long)));
  ADB_CUDA_OK(cudaMemsetAsync(g_dbg, 0, 6 * 256 * sizeof(long long), (cudaStream_t)stream));
  *out = g_dbg;
  return ADB_OK;
}
}

extern "C" int adb_debug_timeline(int64_t* host_out, int32_t count) {
  if (!g_dbg || !host_out || count > 6 * 256) return adbh::fail(ADB_ERR_INVALID, "adb_debug_timeline: no timeline recorded");
  ADB_CUDA_OK(cudaMemcpy(host_out, g_dbg, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost));
  return ADB_OK;
}

extern "C" int adb_conv2d(const adb_conv_desc* d, void* stream) {
  if (d && adbc::roll_eligible(d)) return adbc::conv_roll_launch(d, stream);   // narrow 3x3: rows folded into N
  ConvK P;
  int out_h = 0, out_w = 0, ktot = 0, box_w = 0, box_h = 0;
  int st = build(d, P, out_h, out_w, ktot, box_w, box_h);
  if (st != ADB_OK) return st;
  adbh::DeviceInfo di;
  st = adbh::device_info(&di);
  if (st != ADB_OK) return st;
  if (di.cc_major != 10) return adbh::fail(ADB_ERR_NO_DEVICE, "adb_conv2d: device sm_%d%d is not sm_100", di.cc_major, di.cc_minor);
  P.err_flag = adbh::kernel_err_flag();
  if (d->tune_flags & (4 | 64)) {
    if (!g_dbg) ADB_CUDA_OK(cudaMalloc(&g_dbg, 6 * 256 * sizeof(long long)));
    ADB_CUDA_OK(cudaMemsetAsync(g_dbg, 0, 6 * 256 * sizeof(long long), (cudaStream_t)stream));
    P.dbg = g_dbg;
    P.dbg_detail = (d->tune_flags & 64) ? 1 : 0;
  }

  alignas(64) CUtensorMap tmA0, tmA1, tmB, tmOut;
  const bool s2d_in = d->kind == ADB_CONV_S2;
  st = make_act_tmap(&tmA0, d->src0, d->c0, d->c0_pitch, d->n, d->h_in, d->w_in, s2d_in, P.Ck, box_w, box_h, P.row_bytes);
  if (st != ADB_OK) return st;
  if (d->src1) {
    st = make_act_tmap(&tmA1, d->src1, d->c1, d->c1_pitch, d->n, d->h_in, d->w_in, s2d_in, P.Ck, box_w, box_h, P.row_bytes);
    if (st != ADB_OK) return st;
  } else {
    tmA1 = tmA0;
  }
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)P.ngroups * d->cout_pad};
    uint64_t strides[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {(uint32_t)P.Ck, (uint32_t)(P.BN / P.ncta)};
    st = adbh::make_tmap_bf16(&tmB, d->w_packed, 2, dims, strides, box, P.row_bytes);
    if (st != ADB_OK) return st;
  }
  if (d->epi == ADB_EPI_FEATURE) {
    // one store box per epilogue warp: its 32 tile rows = 32 pixels of one image row, or 32/TW whole rows of a narrow tile
    const int qw = std::min(P.TW, 32), qh = 32 / qw;
    st = make_act_tmap(&tmOut, d->dst, d->dst_pitch, d->dst_pitch, d->n, out_h, out_w, d->kind == ADB_CONVT_4X4S2, P.Cs, qw, qh, P.Cs * 2);
    if (st != ADB_OK) return st;
  } else {
    tmOut = tmA0;
  }

  const bool pre = d->pre_scale != nullptr;
  const SmemLayout L = smem_layout(P.a_slots, P.a_slot_bytes, P.b_slots, P.b_slot_bytes, P.slab_bytes, P.n_slab_bufs, P.cout_pad,
                                   pre ? P.ctot : 0);
  int smem = (int)L.total + 1024;
  smem = std::max(smem, 120 * 1024);  // one CTA per SM: the CTA owns the SM's TMEM
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, ConvK);
  const bool pair = P.ncta == 2;
  KernelFn fn = nullptr;
  int which = 0;
  if (d->epi == ADB_EPI_FEATURE) {
    const int a = d->act == ADB_ACT_RELU ? 0 : (d->act == ADB_ACT_NONE ? 1 : 2);
    which = a * 2 + (pair ? 1 : 0);
    if (pre) which = 8 + a;
    switch (which) {
      case 8: fn = conv_igemm_kernel<ADB_ACT_RELU, false, ADB_EPI_FEATURE, true>; break;
      case 9: fn = conv_igemm_kernel<ADB_ACT_NONE, false, ADB_EPI_FEATURE, true>; break;
      case 10: fn = conv_igemm_kernel<-1, false, ADB_EPI_FEATURE, true>; break;
      case 0: fn = conv_igemm_kernel<ADB_ACT_RELU, false, ADB_EPI_FEATURE>; break;
      case 1: fn = conv_igemm_kernel<ADB_ACT_RELU, true, ADB_EPI_FEATURE>; break;
      case 2: fn = conv_igemm_kernel<ADB_ACT_NONE, false, ADB_EPI_FEATURE>; break;
      case 3: fn = conv_igemm_kernel<ADB_ACT_NONE, true, ADB_EPI_FEATURE>; break;
      case 4: fn = conv_igemm_kernel<-1, false, ADB_EPI_FEATURE>; break;
      default: fn = conv_igemm_kernel<-1, true, ADB_EPI_FEATURE>; break;
    }
  } else {
    ADB_REQUIRE(!pair, "adb_conv2d: the DOT / IMAGE epilogues run in 1-CTA mode only");
    if (d->epi == ADB_EPI_DOT) { fn = conv_igemm_kernel<-1, false, ADB_EPI_DOT>; which = 6; }
    else { fn = conv_igemm_kernel<-1, false, ADB_EPI_IMAGE>; which = 7; }
  }
  static bool configured[11] = {false, false, false, false, false, false, false, false, false, false, false};
  static int max_pairs[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (!configured[which]) {
    ADB_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
    configured[which] = true;
  }
  const long long max_tiles = (long long)d->n * P.tiles_w * P.tiles_h * P.ngroups * P.n_tiles_n;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(pre ? kThreadsPre : kThreads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  if (pair) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (max_pairs[which] == 0) {   // how many CTA pairs the device can hold at once (each CTA owns a whole SM)
      cfg.gridDim = dim3(2 * (di.sm_count / 2), 1, 1);
      int nclusters = 0;
      ADB_CUDA_OK(cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg));
      ADB_REQUIRE(nclusters >= 1, "adb_conv2d: no CTA pair fits the device");
      max_pairs[which] = std::min(nclusters, di.sm_count / 2);
    }
    cfg.gridDim = dim3(2 * (unsigned)std::min<long long>(max_tiles, max_pairs[which]), 1, 1);
  } else {
    cfg.gridDim = dim3((unsigned)std::min<long long>(max_tiles, di.sm_count), 1, 1);
  }
  ADB_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, tmA0, tmA1, tmB, tmOut, P));
  return ADB_OK;
}

extern "C" int64_t adb_conv2d_stat_slots(const adb_conv_desc* d) {
  ConvK P;
  int out_h = 0, out_w = 0, ktot = 0, box_w = 0, box_h = 0;
  adb_conv_desc q = *d;
  q.stat_out = nullptr;
  if (adbc::roll_eligible(&q)) return 0;   // the rolling-row kernel is the faster path for this launch: statistics stay a separate pass
  if (d->epi != ADB_EPI_FEATURE || d->kind == ADB_CONVT_4X4S2 || d->pre_scale) return 0;
  if (build(&q, P, out_h, out_w, ktot, box_w, box_h) != ADB_OK) return -1;
  return (int64_t)P.tiles_w * P.tiles_h * P.ngroups * P.ncta * P.MT * 4 * d->n;
}

extern "C" double adb_conv2d_flops(const adb_conv_desc* d) {
  if (!d) return 0.0;
  // true (unpadded) work: 2 * output pixels * kh*kw*cin * cout
  double cin = (double)d->c0 + d->c1;
  if (d->kind == ADB_CONVT_4X4S2) return 2.0 * d->n * (double)d->h_in * d->w_in * 4.0 * 4.0 * cin * d->cout;
  double oh = d->kind == ADB_CONV_S2 ? d->h_in / 2 : d->h_in;
  double ow = d->kind == ADB_CONV_S2 ? d->w_in / 2 : d->w_in;
  return 2.0 * d->n * oh * ow * d->kh * d->kw * cin * d->cout;
}
