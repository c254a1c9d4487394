// f32x2 (FFMA2) helpers shared by the two fused-activation kernels.
#pragma once
#include "common.cuh"

namespace bvg {

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct TapsPacked {
  f32x2 u[6];   // {up[2q], up[2q]}   (x2 gain included)
  f32x2 d[6];   // {down[k], down[k]}, k = 0..5
};

template <bool FAST>
struct SnakePair {
  f32x2 hb, a2, na2hb, nhb;   // fast path constants
  float a0, a1, ib0, ib1;     // accurate path constants
  __device__ __forceinline__ void init(float al0, float al1, float be0, float be1) {
    a0 = expf(al0); a1 = expf(al1);
    ib0 = 1.0f / (expf(be0) + 1e-9f); ib1 = 1.0f / (expf(be1) + 1e-9f);
    hb = pk2(0.5f * ib0, 0.5f * ib1);
    nhb = pk2(-0.5f * ib0, -0.5f * ib1);
    a2 = pk2(2.0f * a0, 2.0f * a1);
    na2hb = pk2(-2.0f * a0 * 0.5f * ib0, -2.0f * a1 * 0.5f * ib1);
  }
  __device__ __forceinline__ f32x2 acc_init() const { return FAST ? hb : pk2(0.f, 0.f); }
  // `u` is the FIR result (plus hb when FAST)
  __device__ __forceinline__ f32x2 apply(f32x2 u) const {
    if (FAST) {
      float z0, z1;
      upk2(fma2(a2, u, na2hb), z0, z1);
      return fma2(nhb, pk2(__cosf(z0), __cosf(z1)), u);
    } else {
      float u0, u1;
      upk2(u, u0, u1);
      const float s0 = sin_mod_pi(u0 * a0), s1 = sin_mod_pi(u1 * a1);
      return pk2(fmaf(ib0 * s0, s0, u0), fmaf(ib1 * s1, s1, u1));
    }
  }
};


// scalar twins of SnakePair (same operations in the same order)
template <bool FAST>
__device__ __forceinline__ float snake_acc_init(float ib) { return FAST ? 0.5f * ib : 0.f; }
template <bool FAST>
__device__ __forceinline__ float snake_apply(float u, float a, float ib) {
  if (FAST) {
    const float hb = 0.5f * ib;
    const float z = fmaf(2.0f * a, u, -2.0f * a * hb);
    return fmaf(-hb, __cosf(z), u);
  } else {
    const float s = sin_mod_pi(u * a);
    return fmaf(ib * s, s, u);
  }
}

static inline void make_taps_packed(TapsPacked* tp, const Taps& taps) {
  for (int q = 0; q < 6; ++q) {
    union { float f; uint32_t u; } a, b;
    a.f = taps.up[2 * q];
    b.f = taps.down[q];
    tp->u[q] = ((unsigned long long)a.u << 32) | a.u;
    tp->d[q] = ((unsigned long long)b.u << 32) | b.u;
  }
}

}  // namespace bvg
