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

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kMaxASlots = 8;
constexpr int kMaxRing = 16;
constexpr int kStripW = 128;

struct RollK {
  int n, n_start;
  const int* n_dev;
  int H, W;
  int strips, SH, segs_h;          // column strips per row, segment height (output rows), segments per strip
  FastDiv fd_strips, fd_segs;
  int Ck, row_bytes;
  int chunks0, chunks1, pitch0, pitch1, c0, ks_last0, ks_last1;
  int CP, R;                       // cout_pad (columns of one ring slot), ring slots
  int a_slots, a_slot_bytes, a_tx_bytes;
  int b_box_bytes, b_bytes_total;
  uint32_t idesc[3];               // N = CP, 2*CP, 3*CP
  int act;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int res_pitch;
  int Cs, n_slabs, stage_bytes;    // epilogue slab channels, slabs per row, staging bytes per epilogue warp
  int out_c_off;
  int* err_flag;
};

struct RollSmem { uint32_t a_off, b_off, stage_off, scale_off, bar_off, total; };

__host__ __device__ inline RollSmem roll_smem(int a_slots, int a_slot_bytes, int b_bytes, int stage_bytes, int cp) {
  RollSmem L;
  uint32_t off = 0;
  L.a_off = off; off += (uint32_t)a_slots * a_slot_bytes;
  L.b_off = off; off += (uint32_t)b_bytes;
  L.stage_off = off; off += 8u * (uint32_t)stage_bytes;
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

// residual pixels of one output row quarter (32 px x Cs channels), lane-transposed for coalesced 16-byte reads
template <int CS16>
__device__ __forceinline__ void roll_load_residual(const RollK& P, int img, int h, int w0, int sl, int ew, int lane,
                                                   uint4 (&q)[CS16 * 2]) {
  constexpr int CPR = CS16 * 2;
  const int ch = sl * (CS16 * 16) + (lane % CPR) * 8;
#pragma unroll
  for (int i = 0; i < CPR; ++i) {
    const int w = w0 + ew * 32 + i * (32 / CPR) + lane / CPR;
    q[i] = (w < P.W) ? __ldg(reinterpret_cast<const uint4*>(P.residual + (((size_t)img * P.H + h) * P.W + w) * P.res_pitch + ch))
                     : make_uint4(0, 0, 0, 0);
  }
}

template <int kAct, int CS16>
__device__ __forceinline__ void roll_epilogue(const RollK& P, const CUtensorMap* tmOut, uint32_t tmem_base, uint32_t bar_full0,
                                              uint32_t bar_empty0, uint32_t sbuf, const float* s_scale, const float* s_shift,
                                              int ew, int half, int lane, int unit, int nunits, int total) {
  constexpr int Cs = CS16 * 16;
  const bool has_res = P.residual != nullptr;
  const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
  uint32_t full_phase = 0;                                  // one phase bit per ring slot
  // every slot starts zeroed and free
  for (int s = half; s < P.R; s += 2) {
    for (int c = 0; c < P.CP; c += 16) tmem_st16_zero(lane_base + (uint32_t)(s * P.CP + c));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty0 + 8u * s);
  }
  uint4 q[CS16 * 2];
#pragma unroll
  for (int i = 0; i < CS16 * 2; ++i) q[i] = make_uint4(0, 0, 0, 0);
  for (int t = unit; t < total; t += nunits) {
    const Seg sg = decode_seg(P, t);
    for (int i = half; i < sg.rows; i += 2) {
      const int h = sg.h0 + i;
      const int slot = i % P.R;
      if (has_res) roll_load_residual<CS16>(P, sg.img, h, sg.w0, 0, ew, lane, q);   // independent of the accumulator
      mbar_wait(bar_full0 + 8u * slot, (full_phase >> slot) & 1u, P.err_flag, 4);
      full_phase ^= 1u << slot;
      tc_fence_after();
      for (int sl = 0; sl < P.n_slabs; ++sl) {
        if (sl > 0 && has_res) roll_load_residual<CS16>(P, sg.img, h, sg.w0, sl, ew, lane, q);
        if (lane == 0) tma_store_wait_read<0>();             // the previous store has finished reading the staging buffer
        __syncwarp();
        compute_slab<kAct, CS16>(lane_base + (uint32_t)(slot * P.CP + sl * Cs), q, has_res, s_scale + sl * Cs, s_shift + sl * Cs,
                                 sbuf, lane, P.act);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_5d(tmOut, sbuf, P.out_c_off + sl * Cs, sg.w0 + ew * 32, 0, h, sg.img);
          tma_store_commit();
        }
      }
      // hand the slot back zeroed: every MMA accumulates
      for (int c = 0; c < P.CP; c += 16) tmem_st16_zero(lane_base + (uint32_t)(slot * P.CP + c));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bar_empty0 + 8u * slot);
    }
  }
  if (lane == 0) tma_store_wait_all<0>();
}

template <int kAct>
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
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < P.CP; i += 256) { s_scale[i] = P.scale[i]; s_shift[i] = P.shift[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================================================= A producer: one 130-pixel row box per (input row, channel chunk)
    int slot = 0; uint32_t phase = 0;
    for (int t = unit; t < total; t += nunits) {
      const Seg sg = decode_seg(P, t);
      const int jb = max(sg.h0 - 1, 0), je = min(sg.h0 + sg.rows, P.H - 1);
      for (int j = jb; j <= je; ++j) {
        for (int c = 0; c < nchunks; ++c) {
          const bool s1 = c >= P.chunks0;
          const int coff = (s1 ? c - P.chunks0 : c) * P.Ck;
          mbar_wait(emptyA(slot), phase ^ 1u, P.err_flag, 1);
          if (elect_one()) {
            mbar_expect_tx(fullA(slot), (uint32_t)P.a_tx_bytes);
            tma_load_5d(a_base + (uint32_t)slot * P.a_slot_bytes, s1 ? &tmA1 : &tmA0, fullA(slot), coff, sg.w0 - 1, 0, j, sg.img);
          }
          __syncwarp();
          if (++slot == P.a_slots) { slot = 0; phase ^= 1u; }
        }
      }
    }
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
    int sa = 0; uint32_t pa = 0;
    uint32_t empty_phase = 0;                                // one phase bit per ring slot
    const int ksteps_full = P.Ck / 16;
    const uint64_t desc_hi = make_kmajor_desc(0, P.row_bytes);
    if (total > 0) mbar_wait(fullB, 0, P.err_flag, 6);
    for (int t = unit; t < total; t += nunits) {
      const Seg sg = decode_seg(P, t);
      const int jb = max(sg.h0 - 1, 0), je = min(sg.h0 + sg.rows, P.H - 1);
      const int h_last = sg.h0 + sg.rows - 1;
      int next_new = sg.h0, next_commit = sg.h0;
      for (int j = jb; j <= je; ++j) {
        const int hi = min(j + 1, h_last), lo = max(j - 1, sg.h0);
        // output rows entering the window need their (zeroed) ring slot back from the epilogue
        for (; next_new <= hi; ++next_new) {
          const int slot = (next_new - sg.h0) % P.R;
          mbar_wait(empty0 + 8u * slot, (empty_phase >> slot) & 1u, P.err_flag, 2);
          empty_phase ^= 1u << slot;
        }
        tc_fence_after();
        // the window's slots are adjacent except across the ring wrap: one or two MMA pieces
        const int slot_lo = (lo - sg.h0) % P.R;
        const int nrows = hi - lo + 1;
        const int n0 = min(nrows, P.R - slot_lo), n1 = nrows - n0;
        const int blk0 = lo - (j - 1);                       // weight column block of the first window row
        const int done_to = (j == je) ? h_last : j - 1;      // output rows complete after this input row
        for (int c = 0; c < nchunks; ++c) {
          const int ksteps = c == P.chunks0 - 1 ? P.ks_last0 : (c == nchunks - 1 ? P.ks_last1 : ksteps_full);
          mbar_wait(fullA(sa), pa, P.err_flag, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_slot = a_base + (uint32_t)sa * P.a_slot_bytes;
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
              const uint32_t b_box = b_base + (uint32_t)(s * nchunks + c) * P.b_box_bytes;
              const uint64_t a0 = desc_hi | (uint64_t)(((a_slot + (uint32_t)(s * P.row_bytes)) & 0x3FFFFu) >> 4);
              {
                const uint64_t b0 = desc_hi | (uint64_t)(((b_box + (uint32_t)(blk0 * P.CP * P.row_bytes)) & 0x3FFFFu) >> 4);
                const uint32_t d = tmem_base + (uint32_t)(slot_lo * P.CP);
                const uint32_t id = P.idesc[n0 - 1];
                for (int kk = 0; kk < ksteps; ++kk) umma_bf16(d, a0 + (uint64_t)(kk * 2), b0 + (uint64_t)(kk * 2), id, 1u);
              }
              if (n1 > 0) {
                const uint64_t b1 = desc_hi | (uint64_t)(((b_box + (uint32_t)((blk0 + n0) * P.CP * P.row_bytes)) & 0x3FFFFu) >> 4);
                const uint32_t id = P.idesc[n1 - 1];
                for (int kk = 0; kk < ksteps; ++kk) umma_bf16(tmem_base, a0 + (uint64_t)(kk * 2), b1 + (uint64_t)(kk * 2), id, 1u);
              }
            }
            umma_commit(emptyA(sa));
            if (c == nchunks - 1)
              for (int h = next_commit; h <= done_to; ++h) umma_commit(full0 + 8u * ((h - sg.h0) % P.R));
          }
          __syncwarp();
          if (++sa == P.a_slots) { sa = 0; pa ^= 1u; }
        }
        next_commit = max(next_commit, done_to + 1);
      }
    }
  } else if (warp >= kEpiWarp0) {
    const int ewi = warp - kEpiWarp0;
    const int ew = ewi & 3, half = ewi >> 2;
    const uint32_t sbuf = base + L.stage_off + (uint32_t)ewi * (uint32_t)P.stage_bytes;
    if (P.Cs == 64) roll_epilogue<kAct, 4>(P, &tmOut, tmem_base, full0, empty0, sbuf, s_scale, s_shift, ew, half, lane, unit, nunits, total);
    else if (P.Cs == 32) roll_epilogue<kAct, 2>(P, &tmOut, tmem_base, full0, empty0, sbuf, s_scale, s_shift, ew, half, lane, unit, nunits, total);
    else roll_epilogue<kAct, 1>(P, &tmOut, tmem_base, full0, empty0, sbuf, s_scale, s_shift, ew, half, lane, unit, nunits, total);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512u);
}

}  // namespace

namespace adbc {

// Whether adb_conv2d should take the rolling-row kernel for this descriptor (w_fold given by the caller).
bool roll_eligible(const adb_conv_desc* d) {
  if (!d->w_fold || (d->tune_flags & 512)) return false;
  if (d->kind != ADB_CONV_S1 || d->kh != 3 || d->kw != 3 || d->pad != 1 || d->epi != ADB_EPI_FEATURE || d->pre_scale) return false;
  if (d->cout_pad % 32 != 0 || 3 * d->cout_pad > 256) return false;
  if (d->w_in < kStripW) return false;
  if (d->c0 % 16 || d->c1 % 16) return false;
  // the resident filter + three operand slots + staging must fit shared memory
  int Ck = pick_chunk(d->c0);
  if (d->src1) Ck = std::min(Ck, pick_chunk(d->c1));
  if (Ck < 64 && (d->c0 >= 48 || d->c1 >= 48)) Ck = 64;
  const int chunks = (d->c0 + Ck - 1) / Ck + (d->c1 + Ck - 1) / Ck;
  const int b_bytes = 9 * d->cout_pad * Ck * 2 * chunks;
  const int a_slot = round_up(130 * Ck * 2, 1024);
  const int cs = pick_chunk(d->cout_pad);
  return b_bytes + 3 * a_slot + 8 * 32 * cs * 2 + 4096 <= 226 * 1024;
}

int conv_roll_launch(const adb_conv_desc* d, void* stream) {
  RollK P;
  memset(&P, 0, sizeof(P));
  ADB_REQUIRE(d->src0 && d->c0 > 0 && d->c0_pitch >= d->c0 && d->c0_pitch % 8 == 0, "adb_conv2d(roll): bad src0");
  ADB_REQUIRE((d->src1 == nullptr) == (d->c1 == 0), "adb_conv2d(roll): src1/c1 mismatch");
  ADB_REQUIRE(d->n > 0 && d->h_in > 0 && d->w_in >= kStripW, "adb_conv2d(roll): bad n/h/w");
  ADB_REQUIRE(d->dst && d->dst_pitch % 8 == 0 && d->dst_c_off >= 0 && d->dst_c_off % 8 == 0 && d->dst_c_off + d->cout_pad <= d->dst_pitch,
              "adb_conv2d(roll): dst pitch %d cannot hold channels [%d, %d)", d->dst_pitch, d->dst_c_off, d->dst_c_off + d->cout_pad);
  if (d->residual) ADB_REQUIRE(d->res_pitch >= d->cout_pad && d->res_pitch % 8 == 0, "adb_conv2d(roll): residual pitch %d too small", d->res_pitch);
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
  P.CP = d->cout_pad; P.R = 512 / P.CP;
  ADB_REQUIRE(P.R >= 6 && P.R <= kMaxRing && P.R % 2 == 0, "adb_conv2d(roll): ring of %d rows unsupported", P.R);
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
  P.a_tx_bytes = 130 * P.row_bytes;
  P.a_slot_bytes = round_up(P.a_tx_bytes, 1024);
  P.b_box_bytes = 3 * P.CP * P.row_bytes;
  P.b_bytes_total = 3 * nchunks * P.b_box_bytes;
  P.Cs = pick_chunk(P.CP); P.n_slabs = P.CP / P.Cs; P.stage_bytes = 32 * P.Cs * 2;
  const int budget = di.max_smem_optin - 1024;
  const RollSmem fixed = roll_smem(0, 0, P.b_bytes_total, P.stage_bytes, P.CP);
  int a_slots = std::min(kMaxASlots, (budget - (int)fixed.total) / P.a_slot_bytes);
  ADB_REQUIRE(a_slots >= 3, "adb_conv2d(roll): pipeline does not fit shared memory");
  if (d->tune_stages > 0) a_slots = std::min(a_slots, std::max(2, d->tune_stages));
  P.a_slots = a_slots;
  P.act = d->act; P.scale = d->scale; P.shift = d->shift;
  P.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual); P.res_pitch = d->res_pitch;
  P.out_c_off = d->dst_c_off;
  P.err_flag = adbh::kernel_err_flag();

  alignas(64) CUtensorMap tmA0, tmA1, tmB, tmOut;
  st = make_act_tmap(&tmA0, d->src0, d->c0, d->c0_pitch, d->n, P.H, P.W, false, Ck, 130, 1, P.row_bytes);
  if (st != ADB_OK) return st;
  if (d->src1) {
    st = make_act_tmap(&tmA1, d->src1, d->c1, d->c1_pitch, d->n, P.H, P.W, false, Ck, 130, 1, P.row_bytes);
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
  st = make_act_tmap(&tmOut, d->dst, d->dst_pitch, d->dst_pitch, d->n, P.H, P.W, false, P.Cs, 32, 1, P.Cs * 2);
  if (st != ADB_OK) return st;

  const RollSmem L = roll_smem(P.a_slots, P.a_slot_bytes, P.b_bytes_total, P.stage_bytes, P.CP);
  int smem = std::max((int)L.total + 1024, 120 * 1024);     // one CTA per SM: the CTA owns the SM's TMEM
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, RollK);
  const int which = d->act == ADB_ACT_RELU ? 0 : (d->act == ADB_ACT_NONE ? 1 : 2);
  KernelFn fn = which == 0 ? conv_roll_kernel<ADB_ACT_RELU> : (which == 1 ? conv_roll_kernel<ADB_ACT_NONE> : conv_roll_kernel<-1>);
  static bool configured[3] = {false, false, false};
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
