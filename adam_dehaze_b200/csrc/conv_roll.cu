// conv_roll.cu — "rolling-row" 3x3 stride-1 convolution for narrow outputs (3*cout_pad <= 256) on sm_100a.
//
// Why a second conv kernel.  An SS-mode tcgen05.mma re-reads its whole 128-row A slice (4 KB per K step) from shared memory
// whatever N is, so the tap-by-tap implicit GEMM of conv_igemm.cu idles the tensor pipe behind the A operand once
// N = cout <= 64 (40-81 cycles per M128xN32xK16 MMA against 16 of tensor work).  Here the three filter ROWS
// are folded into N instead:
//      E_j[w, (r, co)] = sum_{s, ci} X[j, w + s - 1, ci] * W[co, ci, r, s]          (one input row j, N = 3*cout)
//      out[h, w, co]   = sum_r E_{h + r - 1}[w, (r, co)]
// A CTA walks a 128-pixel-wide column strip top to bottom, one INPUT row per step.  TMEM holds a ring of R output rows
// (R * cout_pad = 512 columns); the MMAs of input row j accumulate straight into the three adjacent ring slots of output
// rows j-1, j, j+1 (column block r = 2, 1, 0 of the weight box), so the sum over r happens inside the accumulator: a
// third of the MMAs and A-operand reads of the tap-by-tap form, no halo rows re-computed between steps, and every input
// row is loaded once.  An output row is complete after input row h+1; the epilogue reads it, stores it, writes zeros
// back (all MMAs run with accumulate = 1) and returns the slot.  The whole filter stays resident in shared memory.
//
// Roles (384 threads, persistent grid of one CTA per SM): warp 0 = A producer (one {Ck, 130 px} TMA box per input row and
// channel chunk, zero-filled outside the image), warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 = weight loader (once),
// warps 4-11 = epilogue (two sets of four lane-quarter warps taking alternate output rows).
//
// Reference arithmetic replaced: the 3x3 nn.Conv2d + BatchNorm2d(eval) + ReLU (+ residual) of ConvBlock / ResidualBlock,
// /root/reference/models/dehazing/base_model.py:4-41 (Light 32->32, Medium 64->64), and torchvision DenseNet's 128->32 conv2.
#include "adb_ptx.cuh"
#include "adb_host.h"
#include "conv_common.cuh"
#include <algorithm>

namespace {

using namespace adb;
using namespace adbc;

constexpr int kEpiWarp0 = 4;
constexpr int kEpiSets = 3;                         // epilogue warp sets (4 lane-quarter warps each) taking every third output row
constexpr int kEpiWarps = 4 * kEpiSets;
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);   // 512
constexpr int kMaxASlots = 8;
constexpr int kMaxRing = 16;
constexpr int kStripW = 128;
constexpr int kMaxG = 4;

struct RollK {
  int n, n_start;
  const int* n_dev;
  int H, W;
  int strips, SH, segs_h;          // column strips per row, segment height (output rows), segments per strip
  FastDiv fd_strips, fd_segs;
  int Ck, row_bytes;
  int chunks0, chunks1, pitch0, pitch1, c0, ks_last0, ks_last1;
  int CP, R, logR;                 // cout_pad (columns of one ring slot), ring slots (a power of two)
  int a_slots, a_slot_bytes, a_tx_bytes;
  int G, a_row_bytes;              // input rows per operand box (one barrier round trip per G rows), bytes of one row in it
  int b_box_bytes, b_bytes_total;
  uint32_t idesc[3];               // N = CP, 2*CP, 3*CP
  int act;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int res_pitch;
  int Cs, n_slabs, stage_bytes;    // epilogue slab channels, slabs per row, staging bytes per epilogue warp
  int out_c_off;
  // DOT / IMAGE heads (cout_pad 16 or 32): same meaning as in conv_igemm.cu
  int epi;
  const float* dot_w; float dot_b; float* dot_out;
  int img_mode;
  const float* img_x; float* img_out; const int* img_index; const float* img_guidance; const float* img_alpha;
  int* err_flag;
  long long* dbg;                  // tune_flags bit 2: wait-cycle statistics of CTA `dbg_cta` (see tools/roll_stats.py)
  int dbg_cta;
};

// stall accounting: `acc += cycles spent in the statement` for the debug CTA only
#define ROLL_T0() const long long _t0 = dbg_on ? clock64() : 0
#define ROLL_ACC(slot) do { if (dbg_on) dbg_acc[slot] += clock64() - _t0; } while (0)

struct RollSmem { uint32_t a_off, b_off, stage_off, scale_off, bar_off, total; };

__host__ __device__ inline RollSmem roll_smem(int a_slots, int a_slot_bytes, int b_bytes, int stage_bytes, int cp) {
  RollSmem L;
  uint32_t off = 0;
  L.a_off = off; off += (uint32_t)a_slots * a_slot_bytes;
  L.b_off = off; off += (uint32_t)b_bytes;
  L.stage_off = off; off += (uint32_t)kEpiWarps * (uint32_t)stage_bytes;
  L.scale_off = off; off += (uint32_t)cp * 8;
  off = (off + 15u) & ~15u;
  L.bar_off = off; off += 8u * (2 * kMaxASlots + 1 + 2 * kMaxRing) + 16;
  L.total = off;
  return L;
}

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
      ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] += A[smem] * B[smem] with the descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_acc(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
      : "memory");
}

struct RollIssue {
  uint32_t hi;          // descriptor high word (SBO, version, swizzle mode)
  uint32_t tap;         // A start-address step per column tap (one pixel row), 16-byte units
  uint32_t sstep;       // B start-address step per column tap (nchunks weight boxes)
};
struct RollRow {        // one input row's window of output rows
  uint32_t d0, id0;     // first piece: TMEM column base, instruction descriptor
  uint32_t id1;         // second piece (across the ring wrap, at TMEM column 0); 0 = none
  uint32_t b0, b1;      // weight descriptor low words of the two pieces (chunk 0, tap 0)
};

// all MMAs of one (input row, channel chunk): 3 column taps x kKs K steps (x 2 pieces across the ring wrap)
template <int kKs, bool kTwo>
__device__ __forceinline__ void roll_issue(const RollIssue& I, const RollRow& r, uint32_t a_lo, uint32_t cb, uint32_t d1) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
#pragma unroll
    for (int kk = 0; kk < kKs; ++kk) {
      const uint32_t a = a_lo + (uint32_t)s * I.tap + (uint32_t)(kk * 2);
      umma_acc(r.d0, a, r.b0 + cb + (uint32_t)s * I.sstep + (uint32_t)(kk * 2), I.hi, r.id0);
      if (kTwo) umma_acc(d1, a, r.b1 + cb + (uint32_t)s * I.sstep + (uint32_t)(kk * 2), I.hi, r.id1);
    }
  }
}

// the rows of one operand box (row g lives a_row_step further on), one straight-line MMA stream per row
template <int kKs>
__device__ __forceinline__ void roll_issue_group(const RollIssue& I, const RollRow (&rr)[kMaxG], int ng, uint32_t a_lo,
                                                 uint32_t a_row_step, uint32_t cb, uint32_t d1) {
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
    if (g < ng) {
      if (rr[g].id1 == 0u) roll_issue<kKs, false>(I, rr[g], a_lo + (uint32_t)g * a_row_step, cb, d1);
      else roll_issue<kKs, true>(I, rr[g], a_lo + (uint32_t)g * a_row_step, cb, d1);
    }
  }
}

struct Seg { int img, w0, h0, rows; };

__device__ __forceinline__ Seg decode_seg(const RollK& P, int t) {
  Seg s;
  uint32_t q = (uint32_t)t, r;
  fast_divmod(q, P.fd_strips, q, r); s.w0 = (int)r * kStripW;
  fast_divmod(q, P.fd_segs, q, r);   s.h0 = (int)r * P.SH;
  s.img = (int)q;
  s.rows = min(P.SH, P.H - s.h0);
  return s;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
      ::"r"(taddr), "r"(z) : "memory");
}

// Epilogue of the rolling-row kernel.  Output rows are numbered g = 0, 1, 2, ... in the order this CTA produces them
// (continuing across its segments); row g lives in ring slot g & (R-1) and its barriers are in phase (g >> logR) & 1, so
// neither side keeps per-slot state.  Set `set` of kEpiSets takes the rows with g % kEpiSets == set; a warp handles the
// 32 pixels of its TMEM lane quarter, 32 channels at a time: TMEM -> registers, slot zeroed and returned at once (every
// MMA accumulates), then affine (+ residual) / activation -> bf16 -> swizzled staging -> one TMA store.
template <int kAct, bool kRes>
__device__ __forceinline__ void roll_epilogue(const RollK& P, const CUtensorMap* tmOut, uint32_t tmem_base, uint32_t bar_full0,
                                              uint32_t bar_empty0, uint32_t sbuf, const float* s_scale, const float* s_shift,
                                              int ew, int set, int lane, int unit, int nunits, int total) {
  const bool dbg_on = P.dbg && (int)blockIdx.x == P.dbg_cta && ew == 0 && set == 0;
  long long dbg_acc[5] = {0, 0, 0, 0, 0};     // full wait, staging wait, slab, zero+return, rows
  const long long dbg_start = dbg_on ? clock64() : 0;
  const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
  const uint32_t rmask = (uint32_t)P.R - 1u;
  // every slot starts zeroed and free
  for (int s = set; s < P.R; s += kEpiSets) {
    for (int c = 0; c < P.CP; c += 32) tmem_st32_zero(lane_base + (uint32_t)(s * P.CP + c));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty0 + 8u * s);
  }
  // staging addresses (64-byte rows, SWIZZLE_64B): this lane's own row (4 chunks), and — for the residual bounce — the
  // chunk this lane fetches in load i (row i*8 + lane/4, chunk lane%4: coalesced 16-byte global reads)
  uint32_t sw_row[4], sw_ld[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sw_row[i] = sbuf + swizzle_addr((uint32_t)lane * 64u + (uint32_t)i * 16u, 64u);
    sw_ld[i] = sbuf + swizzle_addr((uint32_t)(i * 8 + lane / 4) * 64u + (uint32_t)(lane % 4) * 16u, 64u);
  }
  const int px = ew * 32 + (lane >> 2);                    // + i*8: tile pixel of residual load i
  const int res_ch = (lane & 3) * 8;
  uint32_t g_base = 0;                                       // g of the current segment's first row
  uint32_t rem = 0;                                          // g_base % kEpiSets
  for (int t = unit; t < total; t += nunits) {
    const Seg sg = decode_seg(P, t);
    const int first = (set + kEpiSets - (int)rem) % kEpiSets;
    for (int i = first; i < sg.rows; i += kEpiSets) {
      const int h = sg.h0 + i;
      const uint32_t g = g_base + (uint32_t)i;
      const uint32_t slot = g & rmask, par = (g >> P.logR) & 1u;
      const __nv_bfloat16* res_row = kRes ? P.residual + (((size_t)sg.img * P.H + h) * P.W + sg.w0) * P.res_pitch + res_ch : nullptr;
      uint4 q[4];
      if (kRes) {                                             // first slab's residual: independent of the accumulator
#pragma unroll
        for (int k = 0; k < 4; ++k)
          q[k] = (sg.w0 + px + k * 8 < P.W) ? __ldg(reinterpret_cast<const uint4*>(res_row + (size_t)(px + k * 8) * P.res_pitch))
                                            : make_uint4(0, 0, 0, 0);
        // Pull into L2 what this warp reads later without a register to park it in: the other slabs of this row (their loads
        // are issued right before use) and the whole next row this warp owns, kEpiSets rows down — those loads then pay an L2
        // round trip instead of a DRAM one.
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (sg.w0 + px + k * 8 < P.W) {
            const __nv_bfloat16* a = res_row + (size_t)(px + k * 8) * P.res_pitch;
            for (int sl = 1; sl < P.n_slabs; ++sl) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + sl * 32));
            if (i + kEpiSets < sg.rows) {
              const __nv_bfloat16* b = a + (size_t)kEpiSets * P.W * P.res_pitch;
              for (int sl = 0; sl < P.n_slabs; ++sl) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + sl * 32));
            }
          }
        }
      }
      { ROLL_T0(); mbar_wait(bar_full0 + 8u * slot, par, P.err_flag, 4); ROLL_ACC(0); }
      if (dbg_on) dbg_acc[4] += 1;
      tc_fence_after();
      for (int sl = 0; sl < P.n_slabs; ++sl) {
        const long long _ts = dbg_on ? clock64() : 0;
        float v[32];
        tmem_ld32(lane_base + slot * (uint32_t)P.CP + (uint32_t)(sl * 32), v);
        if (sl > 0 && kRes) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            q[k] = (sg.w0 + px + k * 8 < P.W) ? __ldg(reinterpret_cast<const uint4*>(res_row + (size_t)(px + k * 8) * P.res_pitch + sl * 32))
                                              : make_uint4(0, 0, 0, 0);
        }
        { ROLL_T0();
          if (lane == 0) tma_store_wait_read<0>();           // the previous store has finished reading the staging buffer
          __syncwarp();
          ROLL_ACC(1); }
        if (kRes) {                                           // bounce the residual through the staging buffer: lane <- its own row
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sw_ld[k]), "r"(q[k].x), "r"(q[k].y), "r"(q[k].z), "r"(q[k].w) : "memory");
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[k].x), "=r"(q[k].y), "=r"(q[k].z), "=r"(q[k].w) : "r"(sw_row[k]) : "memory");
          __syncwarp();
        }
        tmem_ld_wait();
        if (sl == P.n_slabs - 1) {                            // accumulator row is in registers: zero the slot and hand it back
          ROLL_T0();
          for (int c = 0; c < P.CP; c += 32) tmem_st32_zero(lane_base + slot * (uint32_t)P.CP + (uint32_t)c);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_relaxed(bar_empty0 + 8u * slot);
          ROLL_ACC(3);
        }
        const float* sc_ptr = s_scale + sl * 32;
        const float* sh_ptr = s_shift + sl * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                         // 8 channels = one 16-byte chunk of the staging row
          const float4 sa = *reinterpret_cast<const float4*>(sc_ptr + k * 8), sb = *reinterpret_cast<const float4*>(sc_ptr + k * 8 + 4);
          const float4 ha = *reinterpret_cast<const float4*>(sh_ptr + k * 8), hb = *reinterpret_cast<const float4*>(sh_ptr + k * 8 + 4);
          const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
          const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
          const uint32_t qs[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {                       // two channels per instruction (FFMA2 / FADD2): same bits as fmaf / +
            float y0, y1;
            ffma2(y0, y1, v[k * 8 + 2 * e], v[k * 8 + 2 * e + 1], sc[2 * e], sc[2 * e + 1], sh[2 * e], sh[2 * e + 1]);
            if (kRes) fadd2(y0, y1, y0, y1, __uint_as_float(qs[e] << 16), __uint_as_float(qs[e] & 0xffff0000u));
            pk[e] = (kAct == ADB_ACT_RELU) ? pack_bf16x2_relu(y0, y1) : pack_bf16x2(act_t<kAct>(y0, P.act), act_t<kAct>(y1, P.act));
          }
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sw_row[k]), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_5d(tmOut, sbuf, P.out_c_off + sl * 32, sg.w0 + ew * 32, 0, h, sg.img);
          tma_store_commit();
        }
        if (dbg_on) dbg_acc[2] += clock64() - _ts;
      }
    }
    g_base += (uint32_t)sg.rows;
    rem = (rem + (uint32_t)sg.rows) % kEpiSets;
  }
  if (lane == 0) tma_store_wait_all<0>();
  if (dbg_on && lane == 0) {
    for (int i = 0; i < 5; ++i) P.dbg[32 + i] = dbg_acc[i];
    P.dbg[37] = clock64() - dbg_start;
  }
}

// Epilogue of the 3-channel image heads and the 1-channel guidance / transmission heads on the rolling-row schedule: one
// accumulator row of 16 (or 32) columns per pixel, no staging and no TMA store — the result is fp32 NCHW (IMAGE, written
// through the routed-bucket index) or an fp32 [n,h,w] map (DOT), one coalesced 128-byte store per warp and plane.
template <int kEpi>
__device__ __forceinline__ void roll_epilogue_head(const RollK& P, uint32_t tmem_base, uint32_t bar_full0, uint32_t bar_empty0,
                                                   const float* s_scale, const float* s_shift, int ew, int set, int lane, int unit,
                                                   int nunits, int total) {
  const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
  const uint32_t rmask = (uint32_t)P.R - 1u;
  for (int s = set; s < P.R; s += kEpiSets) {
    for (int c = 0; c < P.CP; c += 16) tmem_st16_zero(lane_base + (uint32_t)(s * P.CP + c));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty0 + 8u * s);
  }
  float dotw[32];
  if (kEpi == ADB_EPI_DOT) {
#pragma unroll
    for (int i = 0; i < 32; ++i) dotw[i] = i < P.CP ? __ldg(P.dot_w + i) : 0.f;
  }
  const size_t plane = (size_t)P.H * P.W;
  const float alpha = (kEpi == ADB_EPI_IMAGE && P.img_mode == ADB_IMG_BLEND) ? __ldg(P.img_alpha) : 0.f;
  uint32_t g_base = 0, rem = 0;
  for (int t = unit; t < total; t += nunits) {
    const Seg sg = decode_seg(P, t);
    const int first = (set + kEpiSets - (int)rem) % kEpiSets;
    const int w = sg.w0 + ew * 32 + lane;
    const bool inb = w < P.W;
    size_t img_row = 0;
    if (kEpi == ADB_EPI_IMAGE) {
      const int pos = P.n_start + sg.img;
      img_row = P.img_index ? (size_t)__ldg(P.img_index + pos) : (size_t)pos;
    }
    for (int i = first; i < sg.rows; i += kEpiSets) {
      const int h = sg.h0 + i;
      const uint32_t g = g_base + (uint32_t)i;
      const uint32_t slot = g & rmask, par = (g >> P.logR) & 1u;
      // the hazy pixel / guidance this lane combines with its accumulator row: fetched before the accumulator is ready
      float xin[3] = {0.f, 0.f, 0.f}, gd = 1.f;
      const size_t img_o = img_row * 3 * plane + (size_t)h * P.W + w;
      if (kEpi == ADB_EPI_IMAGE && inb) {
        if (P.img_mode == ADB_IMG_GUIDED) gd = __ldg(P.img_guidance + ((size_t)sg.img * P.H + h) * P.W + w);
#pragma unroll
        for (int c = 0; c < 3; ++c) xin[c] = __ldg(P.img_x + img_o + c * plane);
      }
      mbar_wait(bar_full0 + 8u * slot, par, P.err_flag, 4);
      tc_fence_after();
      float v[32];
      tmem_ld16(lane_base + slot * (uint32_t)P.CP, v);
      if (kEpi == ADB_EPI_DOT && P.CP > 16) tmem_ld16(lane_base + slot * (uint32_t)P.CP + 16u, v + 16);
      tmem_ld_wait();
      for (int c = 0; c < P.CP; c += 16) tmem_st16_zero(lane_base + slot * (uint32_t)P.CP + (uint32_t)c);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bar_empty0 + 8u * slot);
      if (kEpi == ADB_EPI_DOT) {
        float acc = P.dot_b;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < P.CP) acc = fmaf(apply_act(fmaf(v[c], s_scale[c], s_shift[c]), P.act), dotw[c], acc);
        if (inb) P.dot_out[((size_t)sg.img * P.H + h) * P.W + w] = 1.f / (1.f + __expf(-acc));
      } else if (inb) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float y = apply_act(fmaf(v[c], s_scale[c], s_shift[c]), P.act);
          const float o = (P.img_mode == ADB_IMG_BLEND) ? (1.f - alpha) * xin[c] + alpha * y : fminf(fmaxf(xin[c] + y * gd, 0.f), 1.f);
          P.img_out[img_o + c * plane] = o;
        }
      }
    }
    g_base += (uint32_t)sg.rows;
    rem = (rem + (uint32_t)sg.rows) % kEpiSets;
  }
}

template <int kAct, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
conv_roll_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ RollK P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const RollSmem L = roll_smem(P.a_slots, P.a_slot_bytes, P.b_bytes_total, P.stage_bytes, P.CP);
  const uint32_t a_base = base + L.a_off;
  const uint32_t b_base = base + L.b_off;
  float* s_scale = reinterpret_cast<float*>(base_ptr + L.scale_off);
  float* s_shift = s_scale + P.CP;
  const uint32_t bar_base = base + L.bar_off;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (kMaxASlots + s); };
  const uint32_t fullB = bar_base + 8u * (2 * kMaxASlots);
  const uint32_t full0 = bar_base + 8u * (2 * kMaxASlots + 1);
  const uint32_t empty0 = full0 + 8u * kMaxRing;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + L.bar_off + 8u * (2 * kMaxASlots + 1 + 2 * kMaxRing));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit = (int)blockIdx.x, nunits = (int)gridDim.x;

  int n_eff = P.n;
  if (P.n_dev) n_eff = max(0, min(P.n, *P.n_dev - P.n_start));
  const int total = n_eff * P.strips * P.segs_h;
  const int nchunks = P.chunks0 + P.chunks1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.a_slots; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    mbar_init(fullB, 1);
    for (int s = 0; s < P.R; ++s) { mbar_init(full0 + 8u * s, 1); mbar_init(empty0 + 8u * s, 4); }   // 4 lane-quarter warps per slot
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32((const void*)tmem_ptr_smem), 512u);
    tmem_relinquish();
  }
  if (warp >= kEpiWarp0) {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < P.CP; i += 32 * kEpiWarps) { s_scale[i] = P.scale[i]; s_shift[i] = P.shift[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================================================= A producer: one 130-pixel row box per (input row, channel chunk)
    int slot = 0; uint32_t phase = 0;
    const bool dbg_on = P.dbg && (int)blockIdx.x == P.dbg_cta;
    long long dbg_acc[2] = {0, 0};
    const long long dbg_start = dbg_on ? clock64() : 0;
    for (int t = unit; t < total; t += nunits) {
      const Seg sg = decode_seg(P, t);
      const int jb = max(sg.h0 - 1, 0), je = min(sg.h0 + sg.rows, P.H - 1);
      for (int j = jb; j <= je; j += P.G) {               // rows past je in the last group are loaded (or zero-filled) and ignored
        for (int c = 0; c < nchunks; ++c) {
          const bool s1 = c >= P.chunks0;
          const int coff = (s1 ? c - P.chunks0 : c) * P.Ck;
          { ROLL_T0(); mbar_wait(emptyA(slot), phase ^ 1u, P.err_flag, 1); ROLL_ACC(0); }
          dbg_acc[1] += 1;
          if (elect_one()) {
            mbar_expect_tx(fullA(slot), (uint32_t)P.a_tx_bytes);
            tma_load_5d(a_base + (uint32_t)slot * P.a_slot_bytes, s1 ? &tmA1 : &tmA0, fullA(slot), coff, sg.w0 - 1, 0, j, sg.img);
          }
          __syncwarp();
          if (++slot == P.a_slots) { slot = 0; phase ^= 1u; }
        }
      }
    }
    if (dbg_on && lane == 0) { P.dbg[0] = dbg_acc[0]; P.dbg[1] = dbg_acc[1]; P.dbg[2] = clock64() - dbg_start; }
  } else if (warp == 3) {
    // ======================================================= weights: the whole filter, once
    if (total > 0 && elect_one()) {
      mbar_expect_tx(fullB, (uint32_t)P.b_bytes_total);
      for (int s = 0; s < 3; ++s)
        for (int c = 0; c < nchunks; ++c) {
          const int kc = (c >= P.chunks0 ? P.c0 + (c - P.chunks0) * P.Ck : c * P.Ck);
          tma_load_2d(b_base + (uint32_t)(s * nchunks + c) * P.b_box_bytes, &tmB, fullB, kc, s * 3 * P.CP);
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    // One thread feeds MMAs that take ~N/2 cycles each (48 at N = 96), so the issue path has to be a few instructions per
    // MMA: every descriptor is a 32-bit low word (start address | LBO flag, advanced by register adds) paired with a
    // constant high word; ring slots advance by counters and masks; the taps and K steps are unrolled at compile time.
    int sa = 0; uint32_t pa = 0;
    uint32_t g_base = 0;                                     // running output-row number of the segment's first row (see roll_epilogue)
    const int ksteps_full = P.Ck / 16;
    const uint64_t desc_t = make_kmajor_desc(0, P.row_bytes);
    RollIssue I;
    I.hi = (uint32_t)(desc_t >> 32);
    I.tap = (uint32_t)P.row_bytes >> 4;
    I.sstep = (uint32_t)(nchunks * P.b_box_bytes) >> 4;
    const uint32_t lo_flags = (uint32_t)desc_t;              // LBO field
    const uint32_t a_lo0 = ((a_base & 0x3FFFFu) >> 4) | lo_flags, a_step = (uint32_t)P.a_slot_bytes >> 4;
    const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | lo_flags, box_step = (uint32_t)P.b_box_bytes >> 4;
    const uint32_t blk_step = (uint32_t)(P.CP * P.row_bytes) >> 4;
    const uint32_t rmask = (uint32_t)P.R - 1u;
    const uint32_t a_row_step = (uint32_t)P.a_row_bytes >> 4;
    uint32_t a_lo = a_lo0;
    const bool dbg_on = P.dbg && (int)blockIdx.x == P.dbg_cta;
    long long dbg_acc[4] = {0, 0, 0, 0};          // ring wait, fullA wait, issue, input rows
    const long long dbg_start = dbg_on ? clock64() : 0;
    if (total > 0) mbar_wait(fullB, 0, P.err_flag, 6);
    for (int t = unit; t < total; t += nunits) {
      const Seg sg = decode_seg(P, t);
      const int jb = max(sg.h0 - 1, 0), je = min(sg.h0 + sg.rows, P.H - 1);
      const int h_last = sg.h0 + sg.rows - 1;
      int next_new = sg.h0, next_commit = sg.h0;
      const uint32_t g_off = g_base - (uint32_t)sg.h0;       // g of image row h = g_off + h
      for (int j0 = jb; j0 <= je; j0 += P.G) {
        const int ng = min(P.G, je - j0 + 1);                // input rows of this group
        const int j_last = j0 + ng - 1;
        // output rows entering the group's windows need their (zeroed) ring slot back from the epilogue
        const int hi_g = min(j_last + 1, h_last);
        for (; next_new <= hi_g; ++next_new) {
          const uint32_t g = g_off + (uint32_t)next_new;
          ROLL_T0(); mbar_wait(empty0 + 8u * (g & rmask), (g >> P.logR) & 1u, P.err_flag, 2); ROLL_ACC(0);
        }
        if (dbg_on) dbg_acc[3] += ng;
        // per input row: its window of output rows = ring slots, adjacent except across the ring wrap (one or two MMA pieces)
        RollRow rr[kMaxG];
#pragma unroll
        for (int g = 0; g < kMaxG; ++g) {
          const int j = j0 + g;
          const int hi = min(j + 1, h_last), lo = max(j - 1, sg.h0);
          const uint32_t slot_lo = (g_off + (uint32_t)lo) & rmask;
          const int nrows = hi - lo + 1;
          const int n0 = min(nrows, P.R - (int)slot_lo), n1 = nrows - n0;
          const uint32_t blk0 = (uint32_t)(lo - (j - 1));    // weight column block of the first window row
          rr[g].d0 = tmem_base + slot_lo * (uint32_t)P.CP;
          rr[g].id0 = P.idesc[max(n0, 1) - 1];
          rr[g].id1 = n1 > 0 ? P.idesc[n1 - 1] : 0u;
          rr[g].b0 = b_lo0 + blk0 * blk_step;
          rr[g].b1 = rr[g].b0 + (uint32_t)n0 * blk_step;
        }
        const int done_to = (j_last == je) ? h_last : j_last - 1;   // output rows complete after this group
        for (int c = 0; c < nchunks; ++c) {
          const int ksteps = c == P.chunks0 - 1 ? P.ks_last0 : (c == nchunks - 1 ? P.ks_last1 : ksteps_full);
          { ROLL_T0(); mbar_wait(fullA(sa), pa, P.err_flag, 3); ROLL_ACC(1); }
          tc_fence_after();
          const long long _ti = dbg_on ? clock64() : 0;
          if (elect_one()) {
            const uint32_t cb = (uint32_t)c * box_step;
            if (ksteps == 4) roll_issue_group<4>(I, rr, ng, a_lo, a_row_step, cb, tmem_base);
            else if (ksteps == 2) roll_issue_group<2>(I, rr, ng, a_lo, a_row_step, cb, tmem_base);
            else if (ksteps == 3) roll_issue_group<3>(I, rr, ng, a_lo, a_row_step, cb, tmem_base);
            else roll_issue_group<1>(I, rr, ng, a_lo, a_row_step, cb, tmem_base);
            umma_commit(emptyA(sa));
            if (c == nchunks - 1)
              for (int h = next_commit; h <= done_to; ++h) umma_commit(full0 + 8u * ((g_off + (uint32_t)h) & rmask));
          }
          __syncwarp();
          if (dbg_on) dbg_acc[2] += clock64() - _ti;
          a_lo += a_step;
          if (++sa == P.a_slots) { sa = 0; pa ^= 1u; a_lo = a_lo0; }
        }
        next_commit = max(next_commit, done_to + 1);
      }
      g_base += (uint32_t)sg.rows;
    }
    if (dbg_on && lane == 0) { for (int i = 0; i < 4; ++i) P.dbg[16 + i] = dbg_acc[i]; P.dbg[20] = clock64() - dbg_start; }
  } else if (warp >= kEpiWarp0) {
    const int ewi = warp - kEpiWarp0;
    const int ew = ewi & 3, set = ewi >> 2;
    const uint32_t sbuf = base + L.stage_off + (uint32_t)ewi * (uint32_t)P.stage_bytes;
    if (kEpi != ADB_EPI_FEATURE) roll_epilogue_head<kEpi>(P, tmem_base, full0, empty0, s_scale, s_shift, ew, set, lane, unit, nunits, total);
    else if (P.residual) roll_epilogue<kAct, true>(P, &tmOut, tmem_base, full0, empty0, sbuf, s_scale, s_shift, ew, set, lane, unit, nunits, total);
    else roll_epilogue<kAct, false>(P, &tmOut, tmem_base, full0, empty0, sbuf, s_scale, s_shift, ew, set, lane, unit, nunits, total);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512u);
}

}  // namespace

namespace adbc {

int debug_buffer(long long** out, void* stream);   // conv_igemm.cu

// Whether adb_conv2d should take the rolling-row kernel for this descriptor (w_fold given by the caller).
bool roll_eligible(const adb_conv_desc* d) {
  if (!d->w_fold || (d->tune_flags & 512) || d->stat_out) return false;
  if (d->kind != ADB_CONV_S1 || d->kh != 3 || d->kw != 3 || d->pad != 1 || d->pre_scale) return false;
  if (d->epi == ADB_EPI_FEATURE) {
    if (d->cout_pad % 32 != 0 || 3 * d->cout_pad > 256) return false;
  } else if (d->epi == ADB_EPI_IMAGE) {
    if (d->cout_pad != 16 || d->cout != 3) return false;
  } else if (d->epi == ADB_EPI_DOT) {
    if (d->cout_pad != 16 && d->cout_pad != 32) return false;
    if (d->c0 + d->c1 < 32) return false;      // 16 input channels = 32-byte operand rows: measured slower than the tap-by-tap kernel
  } else {
    return false;
  }
  if (d->w_in < kStripW) return false;
  if (d->c0 % 16 || d->c1 % 16) return false;
  // the resident filter + three operand slots + staging must fit shared memory
  int Ck = pick_chunk(d->c0);
  if (d->src1) Ck = std::min(Ck, pick_chunk(d->c1));
  if (Ck < 64 && (d->c0 >= 48 || d->c1 >= 48)) Ck = 64;
  const int chunks = (d->c0 + Ck - 1) / Ck + (d->c1 + Ck - 1) / Ck;
  const int b_bytes = 9 * d->cout_pad * Ck * 2 * chunks;
  const int a_slot = round_up(130 * Ck * 2, 1024);
  return b_bytes + 3 * a_slot + 12 * 32 * 32 * 2 + 4096 <= 226 * 1024;
}

int conv_roll_launch(const adb_conv_desc* d, void* stream) {
  RollK P;
  memset(&P, 0, sizeof(P));
  ADB_REQUIRE(d->src0 && d->c0 > 0 && d->c0_pitch >= d->c0 && d->c0_pitch % 8 == 0, "adb_conv2d(roll): bad src0");
  ADB_REQUIRE((d->src1 == nullptr) == (d->c1 == 0), "adb_conv2d(roll): src1/c1 mismatch");
  ADB_REQUIRE(d->n > 0 && d->h_in > 0 && d->w_in >= kStripW, "adb_conv2d(roll): bad n/h/w");
  if (d->epi == ADB_EPI_FEATURE) {
    ADB_REQUIRE(d->dst && d->dst_pitch % 8 == 0 && d->dst_c_off >= 0 && d->dst_c_off % 8 == 0 && d->dst_c_off + d->cout_pad <= d->dst_pitch,
                "adb_conv2d(roll): dst pitch %d cannot hold channels [%d, %d)", d->dst_pitch, d->dst_c_off, d->dst_c_off + d->cout_pad);
    if (d->residual) ADB_REQUIRE(d->res_pitch >= d->cout_pad && d->res_pitch % 8 == 0, "adb_conv2d(roll): residual pitch %d too small", d->res_pitch);
  } else if (d->epi == ADB_EPI_DOT) {
    ADB_REQUIRE(d->dot_w && d->dot_out, "adb_conv2d(roll): DOT epilogue needs dot_w/dot_out");
  } else {
    ADB_REQUIRE(d->img_x && d->img_out && d->cout == 3, "adb_conv2d(roll): IMAGE epilogue needs img_x/img_out, cout == 3");
    ADB_REQUIRE(d->img_mode != ADB_IMG_GUIDED || d->img_guidance, "adb_conv2d(roll): GUIDED needs img_guidance");
    ADB_REQUIRE(d->img_mode != ADB_IMG_BLEND || d->img_alpha, "adb_conv2d(roll): BLEND needs img_alpha");
  }
  adbh::DeviceInfo di;
  int st = adbh::device_info(&di);
  if (st != ADB_OK) return st;
  if (di.cc_major != 10) return adbh::fail(ADB_ERR_NO_DEVICE, "adb_conv2d: device sm_%d%d is not sm_100", di.cc_major, di.cc_minor);

  int Ck = pick_chunk(d->c0);
  if (d->src1) Ck = std::min(Ck, pick_chunk(d->c1));
  if (Ck < 64 && (d->c0 >= 48 || d->c1 >= 48)) Ck = 64;     // ragged 64-channel chunks (see conv_igemm.cu)
  P.Ck = Ck; P.row_bytes = Ck * 2;
  P.chunks0 = (d->c0 + Ck - 1) / Ck; P.chunks1 = (d->c1 + Ck - 1) / Ck;
  P.ks_last0 = (d->c0 - (P.chunks0 - 1) * Ck) / 16;
  P.ks_last1 = d->c1 ? (d->c1 - (P.chunks1 - 1) * Ck) / 16 : 0;
  P.pitch0 = d->c0_pitch; P.pitch1 = d->src1 ? d->c1_pitch : d->c0_pitch;
  P.c0 = d->c0;
  const int nchunks = P.chunks0 + P.chunks1;
  const int ctot = d->c0 + d->c1;
  P.CP = d->cout_pad; P.R = std::min(512 / P.CP, kMaxRing);
  P.logR = 0;
  while ((1 << P.logR) < P.R) ++P.logR;
  ADB_REQUIRE(P.R >= 8 && P.R <= kMaxRing && (P.R & (P.R - 1)) == 0, "adb_conv2d(roll): ring of %d rows unsupported (power of two)", P.R);
  for (int i = 0; i < 3; ++i) P.idesc[i] = make_idesc_bf16(128u, (uint32_t)((i + 1) * P.CP));
  P.n = d->n; P.n_start = d->n_start; P.n_dev = d->n_dev;
  P.H = d->h_in; P.W = d->w_in;
  P.strips = (P.W + kStripW - 1) / kStripW;
  // segment height: long enough that the two halo rows are cheap, short enough for a few waves of segments per SM
  int SH = 64;
  if (d->tune_mt > 0) SH = d->tune_mt * 8;
  while (SH > 16 && (long long)d->n * P.strips * ((P.H + SH - 1) / SH) < 4LL * di.sm_count) SH >>= 1;
  P.SH = std::min(SH, P.H);
  P.segs_h = (P.H + P.SH - 1) / P.SH;
  P.fd_strips = make_fastdiv(P.strips); P.fd_segs = make_fastdiv(P.segs_h);
  const long long total = (long long)d->n * P.strips * P.segs_h;
  ADB_REQUIRE((unsigned long long)total * (unsigned long long)std::max(P.strips, P.segs_h) < (1ULL << 32),
              "adb_conv2d(roll): %lld segments exceed the decode range; split the batch", total);
  P.a_row_bytes = 130 * P.row_bytes;
  P.b_box_bytes = 3 * P.CP * P.row_bytes;
  P.b_bytes_total = 3 * nchunks * P.b_box_bytes;
  P.Cs = 32; P.n_slabs = P.CP / P.Cs; P.stage_bytes = 32 * P.Cs * 2;      // 32-channel slabs: 12 epilogue warps at <= 128 registers
  const int budget = di.max_smem_optin - 1024;
  const RollSmem fixed = roll_smem(0, 0, P.b_bytes_total, P.stage_bytes, P.CP);
  const int avail = budget - (int)fixed.total;
  // input rows per operand box: the MMA issuer pays one barrier round trip (and the tensor pipe one bubble) per box, so
  // boxes carry as many rows as leave >= 3 boxes (and >= 2 row groups) in flight and half the TMEM ring to the epilogue
  int G = std::min(kMaxG, P.R / 4);
  if (d->tune_acc_stages > 0 && !(d->tune_flags & 4)) G = std::min(G, d->tune_acc_stages);
  while (G > 1 && round_up(G * P.a_row_bytes, 1024) * std::max(3, 2 * nchunks) > avail) G >>= 1;
  P.G = G;
  P.a_tx_bytes = G * P.a_row_bytes;
  P.a_slot_bytes = round_up(P.a_tx_bytes, 1024);
  int a_slots = std::min(kMaxASlots, avail / P.a_slot_bytes);
  ADB_REQUIRE(a_slots >= 3, "adb_conv2d(roll): pipeline does not fit shared memory");
  if (d->tune_stages > 0) a_slots = std::min(a_slots, std::max(2, d->tune_stages));
  P.a_slots = a_slots;
  P.act = d->act; P.scale = d->scale; P.shift = d->shift;
  P.residual = d->epi == ADB_EPI_FEATURE ? reinterpret_cast<const __nv_bfloat16*>(d->residual) : nullptr; P.res_pitch = d->res_pitch;
  P.out_c_off = d->dst_c_off;
  P.epi = d->epi;
  P.dot_w = d->dot_w; P.dot_b = d->dot_b; P.dot_out = d->dot_out;
  P.img_mode = d->img_mode; P.img_x = d->img_x; P.img_out = d->img_out; P.img_index = d->img_index;
  P.img_guidance = d->img_guidance; P.img_alpha = d->img_alpha;
  P.err_flag = adbh::kernel_err_flag();
  if (d->tune_flags & 4) {
    st = debug_buffer(&P.dbg, stream);
    if (st != ADB_OK) return st;
    P.dbg_cta = d->tune_acc_stages > 0 ? d->tune_acc_stages : 0;     // which CTA reports (tune "acc" doubles as the selector)
  }

  alignas(64) CUtensorMap tmA0, tmA1, tmB, tmOut;
  st = make_act_tmap(&tmA0, d->src0, d->c0, d->c0_pitch, d->n, P.H, P.W, false, Ck, 130, P.G, P.row_bytes);
  if (st != ADB_OK) return st;
  if (d->src1) {
    st = make_act_tmap(&tmA1, d->src1, d->c1, d->c1_pitch, d->n, P.H, P.W, false, Ck, 130, P.G, P.row_bytes);
    if (st != ADB_OK) return st;
  } else {
    tmA1 = tmA0;
  }
  {
    uint64_t dims[2] = {(uint64_t)ctot, (uint64_t)9 * P.CP};
    uint64_t strides[1] = {(uint64_t)ctot * 2};
    uint32_t box[2] = {(uint32_t)Ck, (uint32_t)(3 * P.CP)};
    st = adbh::make_tmap_bf16(&tmB, d->w_fold, 2, dims, strides, box, P.row_bytes);
    if (st != ADB_OK) return st;
  }
  if (d->epi == ADB_EPI_FEATURE) {
    st = make_act_tmap(&tmOut, d->dst, d->dst_pitch, d->dst_pitch, d->n, P.H, P.W, false, P.Cs, 32, 1, P.Cs * 2);
    if (st != ADB_OK) return st;
  } else {
    tmOut = tmA0;
  }

  const RollSmem L = roll_smem(P.a_slots, P.a_slot_bytes, P.b_bytes_total, P.stage_bytes, P.CP);
  int smem = std::max((int)L.total + 1024, 120 * 1024);     // one CTA per SM: the CTA owns the SM's TMEM
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, RollK);
  int which = d->act == ADB_ACT_RELU ? 0 : (d->act == ADB_ACT_NONE ? 1 : 2);
  KernelFn fn = which == 0 ? conv_roll_kernel<ADB_ACT_RELU, ADB_EPI_FEATURE>
                           : (which == 1 ? conv_roll_kernel<ADB_ACT_NONE, ADB_EPI_FEATURE> : conv_roll_kernel<-1, ADB_EPI_FEATURE>);
  if (d->epi == ADB_EPI_DOT) { fn = conv_roll_kernel<-1, ADB_EPI_DOT>; which = 3; }
  if (d->epi == ADB_EPI_IMAGE) { fn = conv_roll_kernel<-1, ADB_EPI_IMAGE>; which = 4; }
  static bool configured[5] = {false, false, false, false, false};
  if (!configured[which]) {
    ADB_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
    configured[which] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cfg.gridDim = dim3((unsigned)std::min<long long>(total, di.sm_count), 1, 1);
  ADB_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, tmA0, tmA1, tmB, tmOut, P));
  return ADB_OK;
}

}  // namespace adbc
