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

// TMEM -> registers -> affine / residual / activation -> bf16 -> the warp's swizzled staging buffer (TMA-store source).
template <int kAct, int CS16>
__device__ __forceinline__ void compute_slab(uint32_t taddr, uint4 (&q)[CS16 * 2], bool has_res, const float* sc_ptr,
                                             const float* sh_ptr, uint32_t sbuf, int lane, int act_rt) {
  constexpr uint32_t span = CS16 * 32;
  constexpr int CPR = CS16 * 2;                      // 16-byte chunks per slab row
  float v[CS16 * 16];
#pragma unroll
  for (int c = 0; c < CS16; ++c) tmem_ld16(taddr + (uint32_t)(c * 16), v + c * 16);
  if (has_res) {
    // q holds the residual lane-transposed (load i, lane l = row i*(32/CPR) + l/CPR, chunk l%CPR: coalesced global
    // reads).  Bounce it through the (free) staging buffer so every lane ends up with its own row.
#pragma unroll
    for (int i = 0; i < CPR; ++i) {
      const uint32_t r = (uint32_t)(i * (32 / CPR) + lane / CPR), c = (uint32_t)(lane % CPR);
      const uint32_t a = sbuf + swizzle_addr(r * span + c * 16u, span);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(q[i].x), "r"(q[i].y), "r"(q[i].z), "r"(q[i].w) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < CPR; ++i) {
      const uint32_t a = sbuf + swizzle_addr((uint32_t)lane * span + (uint32_t)i * 16u, span);
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[i].x), "=r"(q[i].y), "=r"(q[i].z), "=r"(q[i].w) : "r"(a) : "memory");
    }
    __syncwarp();
  }
  tmem_ld_wait();
#pragma unroll
  for (int c16 = 0; c16 < CS16; ++c16) {
    const float4* sc4 = reinterpret_cast<const float4*>(sc_ptr + c16 * 16);
    const float4* sh4 = reinterpret_cast<const float4*>(sh_ptr + c16 * 16);
    float sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = sc4[i], b = sh4[i];
      sc[4 * i] = a.x; sc[4 * i + 1] = a.y; sc[4 * i + 2] = a.z; sc[4 * i + 3] = a.w;
      sh[4 * i] = b.x; sh[4 * i + 1] = b.y; sh[4 * i + 2] = b.z; sh[4 * i + 3] = b.w;
    }
    const uint4 q0 = q[2 * c16], q1 = q[2 * c16 + 1];
    const uint32_t qs[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&qs[i]);
      const float y0 = act_t<kAct>(fmaf(v[c16 * 16 + 2 * i], sc[2 * i], sh[2 * i]) + __low2float(b2), act_rt);
      const float y1 = act_t<kAct>(fmaf(v[c16 * 16 + 2 * i + 1], sc[2 * i + 1], sh[2 * i + 1]) + __high2float(b2), act_rt);
      pk[i] = pack_bf16x2(y0, y1);
    }
    const uint32_t row_off = (uint32_t)lane * span + (uint32_t)c16 * 32u;
    const uint32_t a0 = sbuf + swizzle_addr(row_off, span);
    const uint32_t a1 = sbuf + swizzle_addr(row_off + 16u, span);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
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
