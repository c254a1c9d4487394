// Fused anti-aliased activation on channels-last data [B, T, C]:
//   up x2 (6-tap polyphase of the 12-tap kaiser-sinc) -> Snake/SnakeBeta -> down x2 (12-tap)
// One pass over HBM: each thread owns VEC adjacent channels and slides along a
// time segment keeping a 6-sample x window and a 12-sample v window in
// registers (rotating, so no register moves); consecutive threads own
// consecutive channels, so every load/store is a contiguous row segment.
//
// Semantics follow the reference torch operator (not its CUDA kernel, which is
// wrong at the 3 samples next to each edge - SURVEY.md section 2.3):
//   u[2t]   = sum_q up[2q+1] * x[clamp(t+2-q)]      (up[] already holds the x2 gain)
//   u[2t+1] = sum_q up[2q]   * x[clamp(t+3-q)]
//   v[m]    = u[m] + 1/(b+1e-9) * sin(a*u[m])^2
//   y[t]    = sum_k down[k] * v[clamp(2t+k-5, 0, 2T-1)]
// reference: alias_free_activation/torch/{resample.py:29-38,55-58, filter.py:94-101, act.py:25-30},
//            activations.py:107-120.
#include "common.cuh"

namespace bvg {

template <typename T, int VEC>
struct VecIO;
template <>
struct VecIO<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <>
struct VecIO<float, 2> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x;
    v[1] = t.y;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
};
template <>
struct VecIO<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) {
    v[0] = __bfloat162float(*p);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) {
    *p = __float2bfloat16_rn(v[0]);
  }
};
template <>
struct VecIO<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[2]) {
    uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(p));
    v[0] = __uint_as_float(raw << 16);
    v[1] = __uint_as_float(raw & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[2]) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
  }
};

// One time step of the sliding window for VEC channels.  S is the step index
// modulo 6 (compile time after unrolling): slot (S+5)%6 of X receives x[t+5],
// slots (2S+10)%12,(2S+11)%12 of V receive v[2t+5], v[2t+6].
#define BVG_ACT_UP(S, X, taps, a, ib, vo, ve)                              \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {                        \
    float uo = 0.f, ue = 0.f;                                              \
    _Pragma("unroll") for (int q = 0; q < 6; ++q) {                        \
      const float xv = X[((S) + 5 - q) % 6][j];                            \
      uo = fmaf(taps.up[2 * q], xv, uo);                                   \
      ue = fmaf(taps.up[2 * q + 1], xv, ue);                               \
    }                                                                      \
    vo[j] = snake_eval<FAST>(uo, a[j], ib[j]);                             \
    ve[j] = snake_eval<FAST>(ue, a[j], ib[j]);                             \
  }
#define BVG_ACT_DOWN(S, V, taps, y)                                        \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {                        \
    float acc = 0.f;                                                       \
    _Pragma("unroll") for (int k = 0; k < 12; ++k)                         \
      acc = fmaf(taps.down[k], V[(2 * (S) + k) % 12][j], acc);             \
    y[j] = acc;                                                            \
  }

template <typename Tin, typename Tout, int VEC, bool FAST>
__global__ void __launch_bounds__(128)
act1d_cl_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, const float* __restrict__ alpha_log,
                const float* __restrict__ beta_log, const Taps taps, int B, int64_t T, int C, int L,
                int nseg, int64_t nitems) {
  const int P = C / VEC;
  int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= nitems) return;
  const int pair = (int)(item % P);
  const int64_t rest = item / P;
  const int seg = (int)(rest % nseg);
  const int b = (int)(rest / nseg);
  const int c0 = pair * VEC;

  float a[VEC], ib[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    a[j] = expf(__ldg(alpha_log + c0 + j));
    ib[j] = 1.0f / (expf(__ldg(beta_log + c0 + j)) + 1e-9f);
  }

  const Tin* sp = src + (int64_t)b * T * C + c0;
  Tout* dp = dst + (int64_t)b * T * C + c0;
  const int64_t t0 = (int64_t)seg * L;
  const int64_t t1 = (t0 + L < T) ? t0 + L : T;
  const int64_t tlast = T - 1;

  float X[6][VEC];    // slot p holds x at time == p (mod 6) relative to the first step
  float V[12][VEC];   // slot p holds v at index == p (mod 12) relative to the first step

  // steps run over t = t0-5 .. t1-1; step t loads x[t+5] and produces v[2t+5], v[2t+6], y[t].
  if (t0 >= 5 && t0 + L + 4 <= tlast) {
    // ---- interior segment (L = 6n-5): no clamps, no edge fixes, no predicates ----
    const Tin* lp = sp + (t0 - 5) * C;
    Tout* op = dp + t0 * C;
#pragma unroll
    for (int i = 0; i < 5; ++i) { VecIO<Tin, VEC>::load(lp, X[i]); lp += C; }
    float XN[6][VEC];
#pragma unroll
    for (int i = 0; i < 6; ++i) { VecIO<Tin, VEC>::load(lp, XN[i]); lp += C; }
    // first body: 5 warm-up steps (no output) + 1 full step
#pragma unroll
    for (int s = 0; s < 6; ++s) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) X[(s + 5) % 6][j] = XN[s][j];
      float vo[VEC], ve[VEC];
      BVG_ACT_UP(s, X, taps, a, ib, vo, ve)
#pragma unroll
      for (int j = 0; j < VEC; ++j) { V[(2 * s + 10) % 12][j] = vo[j]; V[(2 * s + 11) % 12][j] = ve[j]; }
      if (s == 5) {
        float y[VEC];
        BVG_ACT_DOWN(s, V, taps, y)
        VecIO<Tout, VEC>::store(op, y);
        op += C;
      }
    }
    const int nbody = (L + 5) / 6 - 1;
    for (int it = 0; it < nbody; ++it) {
      // loads of this body: issued up front, consumed step by step
#pragma unroll
      for (int i = 0; i < 6; ++i) { VecIO<Tin, VEC>::load(lp, XN[i]); lp += C; }
#pragma unroll
      for (int s = 0; s < 6; ++s) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) X[(s + 5) % 6][j] = XN[s][j];
        float vo[VEC], ve[VEC], y[VEC];
        BVG_ACT_UP(s, X, taps, a, ib, vo, ve)
#pragma unroll
        for (int j = 0; j < VEC; ++j) { V[(2 * s + 10) % 12][j] = vo[j]; V[(2 * s + 11) % 12][j] = ve[j]; }
        BVG_ACT_DOWN(s, V, taps, y)
        VecIO<Tout, VEC>::store(op, y);
        op += C;
      }
    }
    return;
  }

  // ---- generic segment: touches a sequence end (replicate clamps, v edge rules) or is short ----
  float vend[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) vend[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    int64_t ti = t0 - 5 + i;
    ti = ti < 0 ? 0 : (ti > tlast ? tlast : ti);
    VecIO<Tin, VEC>::load(sp + ti * C, X[i]);
  }
  const int nsteps = (int)(t1 - t0) + 5;
  for (int base = 0; base < nsteps; base += 6) {
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int64_t t = t0 - 5 + base + s;
      int64_t tl = t + 5;
      tl = tl > tlast ? tlast : tl;  // t+5 >= 0 always
      VecIO<Tin, VEC>::load(sp + tl * C, X[(s + 5) % 6]);
      float vo[VEC], ve[VEC];
      BVG_ACT_UP(s, X, taps, a, ib, vo, ve)
      // right edge: v[m >= 2T] := v[2T-1]; v[2T-1] is the odd sample of step T-3
      if (t >= T - 3) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if (t == T - 3) vend[j] = vo[j];
          vo[j] = vend[j];
          ve[j] = vend[j];
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) { V[(2 * s + 10) % 12][j] = vo[j]; V[(2 * s + 11) % 12][j] = ve[j]; }
      // left edge: v[m < 0] := v[0]; with t0 == 0, v[0] is the even sample of the 3rd step
      if (s == 2 && base == 0 && t0 == 0) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float v0 = V[3][j];
          V[10][j] = v0; V[11][j] = v0; V[0][j] = v0; V[1][j] = v0; V[2][j] = v0;
        }
      }
      float y[VEC];
      BVG_ACT_DOWN(s, V, taps, y)
      if (t >= t0 && t < t1) VecIO<Tout, VEC>::store(dp + t * C, y);
    }
  }
}

template <typename Tin, typename Tout, bool FAST>
static int launch_cl(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                     const Taps& taps, int B, int64_t T, int C, cudaStream_t st) {
  // segment length L = 6n-5 keeps the 6-step unrolled body full; pick the longest
  // one that still gives >= ~4 waves of 128-thread blocks on 148 SMs.
  const bool vec2 = (C % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) % (2 * sizeof(Tin))) == 0) &&
                    ((reinterpret_cast<uintptr_t>(dst) % (2 * sizeof(Tout))) == 0);
  const int VECr = vec2 ? 2 : 1;
  const int64_t P = C / VECr;
  static const int kSegLens[] = {253, 127, 61, 31, 13};  // all of the form 6n-5
  const int64_t want_items = 148LL * 16 * 128;
  int L = kSegLens[4];
  for (int i = 0; i < 5; ++i) {
    if ((int64_t)B * ceil_div(T, kSegLens[i]) * P >= want_items) {
      L = kSegLens[i];
      break;
    }
  }
  const int nseg = (int)ceil_div(T, L);
  const int64_t nitems = (int64_t)B * nseg * P;
  const int threads = 128;
  const int64_t blocks = ceil_div(nitems, threads);
  if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d_cl: tensor too large (%lld blocks)", (long long)blocks);
  if (vec2) {
    act1d_cl_kernel<Tin, Tout, 2, FAST><<<(unsigned)blocks, threads, 0, st>>>(
        (Tout*)dst, (const Tin*)src, alpha_log, beta_log, taps, B, T, C, L, nseg, nitems);
  } else {
    act1d_cl_kernel<Tin, Tout, 1, FAST><<<(unsigned)blocks, threads, 0, st>>>(
        (Tout*)dst, (const Tin*)src, alpha_log, beta_log, taps, B, T, C, L, nseg, nitems);
  }
  BVG_LAUNCHED();
  return BVG_OK;
}

// host entry used by the C ABI and by the vocoder plan
int act1d_cl_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                    const Taps& taps, int B, int64_t T, int C, int in_dtype, int out_dtype, bool fast,
                    cudaStream_t st) {
  if (B <= 0 || T <= 0 || C <= 0) return BVG_OK;
  typedef __nv_bfloat16 bf;
  if (in_dtype == BVG_F32 && out_dtype == BVG_F32)
    return fast ? launch_cl<float, float, true>(dst, src, alpha_log, beta_log, taps, B, T, C, st)
                : launch_cl<float, float, false>(dst, src, alpha_log, beta_log, taps, B, T, C, st);
  if (in_dtype == BVG_F32 && out_dtype == BVG_BF16)
    return fast ? launch_cl<float, bf, true>(dst, src, alpha_log, beta_log, taps, B, T, C, st)
                : launch_cl<float, bf, false>(dst, src, alpha_log, beta_log, taps, B, T, C, st);
  if (in_dtype == BVG_BF16 && out_dtype == BVG_BF16)
    return fast ? launch_cl<bf, bf, true>(dst, src, alpha_log, beta_log, taps, B, T, C, st)
                : launch_cl<bf, bf, false>(dst, src, alpha_log, beta_log, taps, B, T, C, st);
  if (in_dtype == BVG_BF16 && out_dtype == BVG_F32)
    return fast ? launch_cl<bf, float, true>(dst, src, alpha_log, beta_log, taps, B, T, C, st)
                : launch_cl<bf, float, false>(dst, src, alpha_log, beta_log, taps, B, T, C, st);
  BVG_FAIL(BVG_EDTYPE, "act1d_cl: unsupported dtype pair (%d -> %d)", in_dtype, out_dtype);
}

}  // namespace bvg
