// conv_common.cuh — pieces shared by the convolution kernels (conv_igemm.cu, conv_roll.cu): tile-decode division, the
// FEATURE epilogue's slab arithmetic, chunking helpers and the NHWC tensor-map builder.
#pragma once
#include "adb_ptx.cuh"
#include "adb_host.h"
#include <algorithm>
#include <math.h>

namespace adbc {

using namespace adb;

// Division by a launch constant as one multiply-high (tile decode runs once per tile in every role's loop).
// q = umulhi(n, floor(2^32/d) + 1) is exact for n * d < 2^32 (checked on the host against the tile count).
struct FastDiv {
  uint32_t d, mul;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.mul = d <= 1 ? 0u : (uint32_t)(((unsigned long long)1 << 32) / (unsigned long long)d) + 1u;
  return f;
}
__device__ __forceinline__ void fast_divmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
  q = f.d == 1 ? n : __umulhi(n, f.mul);
  r = n - q * f.d;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ADB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == ADB_ACT_TANH) return tanhf(v);
  if (act == ADB_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}
// kAct >= 0: compile-time activation (no per-element branches in the hot FEATURE epilogue); kAct < 0: runtime P.act
template <int kAct>
__device__ __forceinline__ float act_t(float v, int act_rt) {
  if (kAct == ADB_ACT_RELU) return fmaxf(v, 0.f);
  if (kAct == ADB_ACT_NONE) return v;
  return apply_act(v, act_rt);
}

// FEATURE epilogue arithmetic of one slab (32 tile rows x CS16*16 channels; lane = row), in three steps so that the caller can
// order them around the staging buffer's availability:
//   slab_tmem_load   issue the TMEM loads of all the slab's columns (one tcgen05.wait::ld later)
//   slab_bounce      kRes: the residual arrives lane-transposed (coalesced global reads, `src`); pass it through the (free)
//                    staging buffer so every lane ends up with its own row (`q`)
//   slab_math        affine (+ residual) + activation -> packed bf16 pairs.  Two channels per instruction: FFMA2 (+ FADD2),
//                    and ReLU rides inside the bf16 conversion (cvt.rn.relu) — bit-identical to the scalar fma / add / max / cvt
//   slab_stage       the packed row -> this lane's swizzled row of the staging buffer (TMA-store source)
template <int NG>
__device__ __forceinline__ void slab_tmem_load(uint32_t taddr, float (&v)[NG * 16]) {
#pragma unroll
  for (int c = 0; c < NG; ++c) tmem_ld16(taddr + (uint32_t)(c * 16), v + c * 16);
}

template <int CS16>
__device__ __forceinline__ void slab_bounce(const uint4 (&src)[CS16 * 2], uint4 (&q)[CS16 * 2], uint32_t sbuf, int lane) {
  constexpr uint32_t span = CS16 * 32;
  constexpr int CPR = CS16 * 2;                      // 16-byte chunks per slab row
#pragma unroll
  for (int i = 0; i < CPR; ++i) {                    // load i, lane l holds row i*(32/CPR) + l/CPR, chunk l%CPR
    const uint32_t r = (uint32_t)(i * (32 / CPR) + lane / CPR), c = (uint32_t)(lane % CPR);
    const uint32_t a = sbuf + swizzle_addr(r * span + c * 16u, span);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(src[i].x), "r"(src[i].y), "r"(src[i].z), "r"(src[i].w) : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < CPR; ++i) {
    const uint32_t a = sbuf + swizzle_addr((uint32_t)lane * span + (uint32_t)i * 16u, span);
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[i].x), "=r"(q[i].y), "=r"(q[i].z), "=r"(q[i].w) : "r"(a) : "memory");
  }
  __syncwarp();
}

// slab_math handles the 16-channel groups G0 .. G0+NG-1 of a CS16-group slab; v holds just those groups' accumulators.
// kStageNow: each group is written to the staging row as soon as it is packed (keeps 8 instead of CS16*8 packed registers
// live — the residual path also holds this and the next slab's residual rows); pk is then scratch.
template <int kAct, int CS16, int G0, int NG, bool kRes, bool kStageNow>
__device__ __forceinline__ void slab_math(const float (&v)[NG * 16], const uint4 (&q)[CS16 * 2], const float* sc_ptr,
                                          const float* sh_ptr, int act_rt, uint32_t (&pk)[CS16 * 8], uint32_t sbuf, int lane) {
  constexpr uint32_t span = CS16 * 32;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c16 = G0 + g;
    const float4* sc4 = reinterpret_cast<const float4*>(sc_ptr + c16 * 16);
    const float4* sh4 = reinterpret_cast<const float4*>(sh_ptr + c16 * 16);
    float sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = sc4[i], b = sh4[i];
      sc[4 * i] = a.x; sc[4 * i + 1] = a.y; sc[4 * i + 2] = a.z; sc[4 * i + 3] = a.w;
      sh[4 * i] = b.x; sh[4 * i + 1] = b.y; sh[4 * i + 2] = b.z; sh[4 * i + 3] = b.w;
    }
    const uint32_t qs[8] = {q[2 * c16].x, q[2 * c16].y, q[2 * c16].z, q[2 * c16].w,
                            q[2 * c16 + 1].x, q[2 * c16 + 1].y, q[2 * c16 + 1].z, q[2 * c16 + 1].w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y0, y1;
      ffma2(y0, y1, v[g * 16 + 2 * i], v[g * 16 + 2 * i + 1], sc[2 * i], sc[2 * i + 1], sh[2 * i], sh[2 * i + 1]);
      if (kRes) fadd2(y0, y1, y0, y1, __uint_as_float(qs[i] << 16), __uint_as_float(qs[i] & 0xffff0000u));   // bf16 pair -> fp32
      if (kAct == ADB_ACT_RELU) pk[c16 * 8 + i] = pack_bf16x2_relu(y0, y1);
      else if (kAct == ADB_ACT_NONE) pk[c16 * 8 + i] = pack_bf16x2(y0, y1);
      else pk[c16 * 8 + i] = pack_bf16x2(apply_act(y0, act_rt), apply_act(y1, act_rt));
    }
    if (kStageNow) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t a = sbuf + swizzle_addr((uint32_t)lane * span + (uint32_t)(c16 * 2 + i) * 16u, span);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(pk[c16 * 8 + 4 * i]), "r"(pk[c16 * 8 + 4 * i + 1]),
                     "r"(pk[c16 * 8 + 4 * i + 2]), "r"(pk[c16 * 8 + 4 * i + 3]) : "memory");
      }
    }
  }
}

template <int CS16>
__device__ __forceinline__ void slab_stage(const uint32_t (&pk)[CS16 * 8], uint32_t sbuf, int lane) {
  constexpr uint32_t span = CS16 * 32;
#pragma unroll
  for (int i = 0; i < CS16 * 2; ++i) {
    const uint32_t a = sbuf + swizzle_addr((uint32_t)lane * span + (uint32_t)i * 16u, span);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(pk[4 * i]), "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
  }
}


inline int pick_chunk(int c) { return (c % 64 == 0) ? 64 : (c % 32 == 0) ? 32 : (c % 16 == 0) ? 16 : 0; }
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int pow2_at_least(int v) { int p = 32; while (p < v) p <<= 1; return p; }

// 5-D view {C, W, P, H, N} of an NHWC bf16 buffer; s2d = space-to-depth (stride-2 read / sub-pixel write) view.
// c_dim = the channel extent a box may touch (plain view only): an input map's own channel count, so that a ragged
// 64-channel box over a narrower source — a channel slice of a wider buffer included — is zero-filled past it instead of
// reading a neighbour's channels or running past the end of the allocation; the pitch for an output map.
inline int make_act_tmap(CUtensorMap* m, const void* base, int c_dim, int pitch, int n, int h, int w, bool s2d, int box_c, int box_w,
                  int box_h, int span) {
  uint64_t dims[5], strides[4];
  uint32_t box[5] = {(uint32_t)box_c, (uint32_t)box_w, 1u, (uint32_t)box_h, 1u};
  const uint64_t px = (uint64_t)pitch * 2;
  if (!s2d) {
    dims[0] = c_dim; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = n;
    strides[0] = px; strides[1] = px * w; strides[2] = px * w; strides[3] = px * w * h;
  } else {
    dims[0] = 2 * (uint64_t)pitch; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = n;
    strides[0] = 2 * px; strides[1] = px * w; strides[2] = 2 * px * w; strides[3] = px * w * h;
  }
  return adbh::make_tmap_bf16(m, base, 5, dims, strides, box, span);
}


}  // namespace adbc
