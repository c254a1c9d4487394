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
#include <cstring>

#include "act_packed.cuh"

namespace bvg {

// 8 pad channels right behind a channel pair (row pitch = channels + 8): exact zeros, so that the zero weight columns of the
// next conv never meet a stale NaN bit pattern.  Issued by the thread that owns the LAST real pair, at a constant offset.
__device__ __forceinline__ void store_pad8(float* p) {
  *(reinterpret_cast<float4*>(p + 2)) = make_float4(0.f, 0.f, 0.f, 0.f);
  *(reinterpret_cast<float4*>(p + 6)) = make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void store_pad8(__nv_bfloat16* p) { *(reinterpret_cast<uint4*>(p + 2)) = make_uint4(0u, 0u, 0u, 0u); }


template <typename T, int VEC>
struct VecIO;
template <>
struct VecIO<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = BVG_LDG(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *(p) = v[0]; }
};
template <>
struct VecIO<float, 2> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
    float2 t = BVG_LDG(reinterpret_cast<const float2*>(p));
    v[0] = t.x;
    v[1] = t.y;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[2]) {
    *(reinterpret_cast<float2*>(p)) = make_float2(v[0], v[1]);
  }
};
template <>
struct VecIO<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) {
    v[0] = ld_mut_f32<__nv_bfloat16>(p);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) {
    *(reinterpret_cast<unsigned short*>(p)) = __bfloat16_as_ushort(__float2bfloat16_rn(v[0]));
  }
};
template <>
struct VecIO<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[2]) {
    uint32_t raw = BVG_LDG(reinterpret_cast<const uint32_t*>(p));
    v[0] = __uint_as_float(raw << 16);
    v[1] = __uint_as_float(raw & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[2]) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[0], v[1]);
    *(reinterpret_cast<unsigned int*>(p)) = *reinterpret_cast<const unsigned int*>(&h2);
  }
};

// One time step of the sliding window for VEC channels.  S is the step index
// modulo 6 (compile time after unrolling): slot (S+5)%6 of X receives x[t+5],
// slots (2S+10)%12,(2S+11)%12 of V receive v[2t+5], v[2t+6].
// The arithmetic (operation order, fma/mul placement) is bit-identical to the packed
// kernel's, so the result does not depend on how a tensor is cut into segments.
#define BVG_ACT_UP(S, X, taps, a, ib, vo, ve)                              \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {                        \
    float uo = snake_acc_init<FAST>(ib[j]), ue = uo;                       \
    _Pragma("unroll") for (int q = 0; q < 6; ++q) {                        \
      const float xv = X[((S) + 5 - q) % 6][j];                            \
      uo = fmaf(taps.up[2 * q], xv, uo);                                   \
      ue = fmaf(taps.up[2 * q + 1], xv, ue);                               \
    }                                                                      \
    vo[j] = snake_apply<FAST>(uo, a[j], ib[j]);                            \
    ve[j] = snake_apply<FAST>(ue, a[j], ib[j]);                            \
  }
#define BVG_ACT_DOWN(S, V, taps, y)                                        \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {                        \
    float acc = taps.down[0] * V[(2 * (S)) % 12][j];                       \
    _Pragma("unroll") for (int k = 1; k < 12; ++k)                         \
      acc = fmaf(taps.down[k], V[(2 * (S) + k) % 12][j], acc);             \
    y[j] = acc;                                                            \
  }

template <typename Tin, typename Tout, int VEC, bool FAST, bool PAD8 = false>
__device__ __forceinline__ void act1d_cl_body(Tout* __restrict__ dst, const Tin* __restrict__ src,
                                              const float* __restrict__ alpha_log, const float* __restrict__ beta_log,
                                              const Taps& taps, int B, int64_t T, int C, int ld, int L, int nseg, int nseg_head,
                                              int64_t tail_start, int64_t nitems, int64_t item) {
  // `nseg` segments of length L per utterance: `nseg_head` segments from row 0 and the rest from row
  // `tail_start` (the rows in between belong to the packed path)
  const int P = C / VEC;
  if (item >= nitems) return;
  const int pair = (int)(item % P);
  const int64_t rest = item / P;
  const int seg = (int)(rest % nseg);
  const int b = (int)(rest / nseg);
  const int c0 = pair * VEC;
  const bool zpad = PAD8 && VEC == 2 && pair == P - 1;   // see store_pad8

  float a[VEC], ib[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    a[j] = expf(__ldg(alpha_log + c0 + j));
    ib[j] = 1.0f / (expf(__ldg(beta_log + c0 + j)) + 1e-9f);
  }

  const Tin* sp = src + (int64_t)b * T * ld + c0;
  Tout* dp = dst + (int64_t)b * T * ld + c0;
  const bool head = seg < nseg_head;
  const int64_t t0 = head ? (int64_t)seg * L : tail_start + (int64_t)(seg - nseg_head) * L;
  const int64_t rend = head ? (tail_start < T ? tail_start : T) : T;   // end of this segment's region
  const int64_t t1 = (t0 + L < rend) ? t0 + L : rend;
  const int64_t tlast = T - 1;

  float X[6][VEC];    // slot p holds x at time == p (mod 6) relative to the first step
  float V[12][VEC];   // slot p holds v at index == p (mod 12) relative to the first step

  // steps run over t = t0-5 .. t1-1; step t loads x[t+5] and produces v[2t+5], v[2t+6], y[t].
  if (t0 >= 5 && t0 + L + 4 <= tlast && t1 - t0 == L) {
    // ---- interior segment (L = 6n-5): no clamps, no edge fixes, no predicates ----
    const Tin* lp = sp + (t0 - 5) * ld;
    Tout* op = dp + t0 * ld;
#pragma unroll
    for (int i = 0; i < 5; ++i) { VecIO<Tin, VEC>::load(lp, X[i]); lp += ld; }
    float XN[6][VEC];
#pragma unroll
    for (int i = 0; i < 6; ++i) { VecIO<Tin, VEC>::load(lp, XN[i]); lp += ld; }
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
        if (zpad) store_pad8(op);
        op += ld;
      }
    }
    const int nbody = (L + 5) / 6 - 1;
    for (int it = 0; it < nbody; ++it) {
      // loads of this body: issued up front, consumed step by step
#pragma unroll
      for (int i = 0; i < 6; ++i) { VecIO<Tin, VEC>::load(lp, XN[i]); lp += ld; }
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
        if (zpad) store_pad8(op);
        op += ld;
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
    VecIO<Tin, VEC>::load(sp + ti * ld, X[i]);
  }
  const int nsteps = (int)(t1 - t0) + 5;
  for (int base = 0; base < nsteps; base += 6) {
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int64_t t = t0 - 5 + base + s;
      int64_t tl = t + 5;
      tl = tl > tlast ? tlast : tl;  // t+5 >= 0 always
      VecIO<Tin, VEC>::load(sp + tl * ld, X[(s + 5) % 6]);
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
      if (t >= t0 && t < t1) {
        VecIO<Tout, VEC>::store(dp + t * ld, y);
        if (zpad) store_pad8(dp + t * ld);
      }
    }
  }
}

template <typename Tin, typename Tout, int VEC, bool FAST>
__global__ void __launch_bounds__(128)
act1d_cl_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, const float* __restrict__ alpha_log,
                const float* __restrict__ beta_log, const Taps taps, int B, int64_t T, int C, int ld, int L,
                int nseg, int nseg_head, int64_t tail_start, int64_t nitems) {
  pdl_launch_dependents();
  pdl_wait();
  act1d_cl_body<Tin, Tout, VEC, FAST>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, L, nseg, nseg_head, tail_start, nitems,
                                      (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// Packed variant (2 channels per thread as one f32x2 lane pair): Blackwell's FFMA2
// (fma.rn.f32x2) does two FMAs per issue slot.  ncu showed the scalar kernel is
// issue-bound (issue-active ~80 %, 45 lane-instructions per element); packing the two
// channel FIRs halves the FMA issue count (-> ~23 per element).
//   taps are passed pre-duplicated {f,f}; by symmetry f[k] = f[11-k] only 6 distinct
//   pairs per filter are needed: up[2q] = U[q], up[2q+1] = U[5-q], down[k] = D[min(k,11-k)].
//   fast snake: v = (u + hb) - hb*cos(2a*u) with the FIR accumulator initialised to hb,
//   so u + hb is free and 2a*u = fma(2a, u + hb, -2a*hb).
template <typename T> struct PairIO;
template <> struct PairIO<float> {
  typedef float2 raw_t;
  static __device__ __forceinline__ raw_t ldraw(const float* p) { return BVG_LDG(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ f32x2 cvt(raw_t r) { return pk2(r.x, r.y); }
  static __device__ __forceinline__ void store(float* p, f32x2 v) {
    float a, b; upk2(v, a, b);
    *(reinterpret_cast<float2*>(p)) = make_float2(a, b);
  }
};
template <> struct PairIO<__nv_bfloat16> {
  typedef uint32_t raw_t;
  static __device__ __forceinline__ raw_t ldraw(const __nv_bfloat16* p) { return BVG_LDG(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ f32x2 cvt(raw_t r) { return pk2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u)); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, f32x2 v) {
    float a, b; upk2(v, a, b);
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    *(reinterpret_cast<unsigned int*>(p)) = *reinterpret_cast<const unsigned int*>(&h2);
  }
};

#define BVG_ACT2_STEP(S, WITH_DOWN)                                                        \
  {                                                                                        \
    f32x2 uo = sn.acc_init(), ue = uo;                                                     \
    _Pragma("unroll") for (int q = 0; q < 6; ++q) {                                        \
      const f32x2 xv = X[((S) + 5 - q) % 6];                                               \
      uo = fma2(tp.u[q], xv, uo);                                                          \
      ue = fma2(tp.u[5 - q], xv, ue);                                                      \
    }                                                                                      \
    V[(2 * (S) + 10) % 12] = sn.apply(uo);                                                 \
    V[(2 * (S) + 11) % 12] = sn.apply(ue);                                                 \
    if (WITH_DOWN) {                                                                       \
      f32x2 acc = mul2(tp.d[0], V[(2 * (S)) % 12]);                                        \
      _Pragma("unroll") for (int k = 1; k < 12; ++k)                                       \
        acc = fma2(tp.d[k < 6 ? k : 11 - k], V[(2 * (S) + k) % 12], acc);                  \
      PairIO<Tout>::store(op, acc);                                                        \
      if (PAD8 && zpad) store_pad8(op);                                                    \
      op += ld;                                                                             \
    }                                                                                      \
  }

// interior segments only (t0 >= 5, t0 + L + 4 <= T - 1, L = 6n-5); edge segments are
// handled by the scalar kernel's generic path in a second (tiny) launch.
// 256-thread blocks: one block (8 warps, <= 96 registers) fits on an SM next to a persistent
// tcgen05 conv CTA of another AMP block (vocoder.cu: blocks of a stage run on separate streams).
#ifndef BVG_ACT_THREADS
#define BVG_ACT_THREADS 128
#endif
constexpr int kPackedThreads = BVG_ACT_THREADS;
#ifndef BVG_ACT_PF
#define BVG_ACT_PF 6           // rows of run-ahead loads per thread (6 or 12)
#endif
#ifndef BVG_ACT_MINBLK
#define BVG_ACT_MINBLK 6       // __launch_bounds__ minimum blocks per SM: 80 registers, 24 warps per SM - measured 10 % faster
                               // than the unconstrained 96-register build (5 blocks); 7-8 blocks spill and give it back
#endif
template <typename Tin, typename Tout, bool FAST, bool PAD8>
__global__ void __launch_bounds__(kPackedThreads, BVG_ACT_MINBLK)
act1d_cl_packed_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, const float* __restrict__ alpha_log,
                       const float* __restrict__ beta_log, const TapsPacked tp, int B, int64_t T, int C, int ld, int L,
                       int nseg_int, int head_len, int64_t nitems, const Taps taps, int main_blocks, int nseg_edge,
                       int nseg_head, int64_t tail_start, int64_t nitems_edge) {
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.x >= main_blocks) {
    // the last blocks of the grid take the sequence ends (short segments, scalar edge-aware path): they run
    // beside the interior blocks instead of as a separate ~10 us launch behind them
    act1d_cl_body<Tin, Tout, 2, FAST, PAD8>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, head_len, nseg_edge, nseg_head, tail_start,
                                      nitems_edge, (int64_t)(blockIdx.x - main_blocks) * blockDim.x + threadIdx.x);
    return;
  }
  const int P = C / 2;
  int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= nitems) return;
  const int pair = (int)(item % P);
  const int64_t rest = item / P;
  const int seg = (int)(rest % nseg_int);
  const int b = (int)(rest / nseg_int);
  const int c0 = pair * 2;
  const bool zpad = PAD8 && pair == P - 1;   // see store_pad8
  SnakePair<FAST> sn;
  sn.init(__ldg(alpha_log + c0), __ldg(alpha_log + c0 + 1), __ldg(beta_log + c0), __ldg(beta_log + c0 + 1));

  const int64_t t0 = head_len + (int64_t)seg * L;   // interior region [head_len, head_len + nseg_int*L)
  const Tin* lp = src + ((int64_t)b * T + (t0 - 5)) * ld + c0;
  Tout* op = dst + ((int64_t)b * T + t0) * ld + c0;

  f32x2 X[6], V[12];
  // Rolling prefetch: R[s] holds the raw row consumed by step s of the current PF-step
  // iteration; right after it is consumed the slot is refilled with the row PF steps ahead, so
  // every load has PF steps (> 1k cycles) to land.  The launcher keeps 12 spare rows behind the
  // interior region, so the run-ahead loads of the last segment stay inside the tensor.
  constexpr int PF = BVG_ACT_PF;
  typename PairIO<Tin>::raw_t R[PF];
#pragma unroll
  for (int i = 0; i < 5; ++i) { X[i] = PairIO<Tin>::cvt(PairIO<Tin>::ldraw(lp)); lp += ld; }
#pragma unroll
  for (int i = 0; i < PF; ++i) { R[i] = PairIO<Tin>::ldraw(lp); lp += ld; }
  const int niter = (L + 5) / PF;   // L = 12m-5
  // first iteration: steps 0..4 are warm-up (no output)
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    X[(s + 5) % 6] = PairIO<Tin>::cvt(R[s]);
    R[s] = PairIO<Tin>::ldraw(lp);
    lp += ld;
    if (s < 5) BVG_ACT2_STEP(s, false) else BVG_ACT2_STEP(s, true)
  }
  for (int it = 1; it < niter; ++it) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      X[(s + 5) % 6] = PairIO<Tin>::cvt(R[s]);
      R[s] = PairIO<Tin>::ldraw(lp);
      lp += ld;
      BVG_ACT2_STEP(s, true)
    }
  }
}

// Launch plan: rows [kEdge, kEdge + n_int*L) of every utterance ("interior": no clamps, no edge
// rules) go to the packed FFMA2 kernel in segments of L = 6n-5 rows; the head [0, kEdge) and the
// tail go to the scalar edge-aware kernel in SHORT segments (13 rows) so that the few edge
// threads are not a serial tail behind the main kernel.
template <typename Tin, typename Tout, bool FAST>
static int launch_cl(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                     const Taps& taps, int B, int64_t T, int C, int ld, cudaStream_t st) {
  // C channels of every row are processed; rows are ld >= C elements apart (pad channels are neither read nor written)
  const bool vec2 = (C % 2 == 0) && (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) % (2 * sizeof(Tin))) == 0) &&
                    ((reinterpret_cast<uintptr_t>(dst) % (2 * sizeof(Tout))) == 0);
  const int threads = 128;
  constexpr int kEdge = 13;   // 6n-5, >= 5
  if (!vec2) {
    // odd channel count / unaligned: scalar kernel over everything
    if (ld > C) BVG_FAIL(BVG_EINVAL, "act1d_cl: a row pitch wider than the channel count needs even channels and aligned rows");
    static const int kSegLens[] = {253, 127, 61, 31, 13};
    int L = kSegLens[4];
    for (int i = 0; i < 5; ++i)
      if ((int64_t)B * ceil_div(T, kSegLens[i]) * C >= 148LL * 16 * 128) { L = kSegLens[i]; break; }
    const int nseg = (int)ceil_div(T, L);
    const int64_t nitems = (int64_t)B * nseg * C;
    const int64_t blocks = ceil_div(nitems, threads);
    if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d_cl: tensor too large (%lld blocks)", (long long)blocks);
    launch_pdl(act1d_cl_kernel<Tin, Tout, 1, FAST>, (unsigned)blocks, threads, 0, st,
        (Tout*)dst, (const Tin*)src, alpha_log, beta_log, taps, B, T, C, ld, L, nseg, nseg, T, nitems);
    BVG_LAUNCHED();
    return BVG_OK;
  }
  const int64_t P = C / 2;
  static const int kSegLens[] = {247, 127, 55, 31, 19};  // all of the form 12m-5
  const int64_t want_items = 148LL * 16 * 128;
  int L = kSegLens[4];
  for (int i = 0; i < 5; ++i) {
    if ((int64_t)B * ceil_div(T, kSegLens[i]) * P >= want_items) {
      L = kSegLens[i];
      break;
    }
  }
  // one short utterance: the kernel's duration is a thread's serial walk (L + 5 steps), not its throughput - 7-row segments
  // (12 steps for 7 outputs) halve it while the grid is still far from filling the machine
  if (L == kSegLens[4] && (int64_t)B * ceil_div(T, kSegLens[4]) * P < want_items / 4) L = 7;
  // interior segments: kEdge + s*L + L + 4 + 12 (prefetch run-ahead) <= T - 1
  int64_t n_int = (T - 5 - 12 - kEdge) / L;
  if (n_int < 0) n_int = 0;
  // edges: head [0, kEdge) (or everything when there is no interior) and tail [tail_start, T)
  const int64_t tail_start = n_int > 0 ? kEdge + n_int * L : T;
  const int64_t head_end = n_int > 0 ? kEdge : T;
  const int nseg_head = (int)ceil_div(head_end, kEdge);
  const int nseg_tail = (int)ceil_div(T - tail_start, kEdge);
  const int nseg = nseg_head + nseg_tail;
  const int64_t nitems_edge = (int64_t)B * nseg * P;
  if (n_int > 0) {
    TapsPacked tp;
    make_taps_packed(&tp, taps);
    const int64_t nitems = (int64_t)B * n_int * P;
    const int64_t main_blocks = ceil_div(nitems, kPackedThreads);
    const int64_t blocks = main_blocks + ceil_div(nitems_edge, kPackedThreads);
    if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d_cl: tensor too large (%lld blocks)", (long long)blocks);
    if (ld > C)
      launch_pdl(act1d_cl_packed_kernel<Tin, Tout, FAST, true>, (unsigned)blocks, kPackedThreads, 0, st,
          (Tout*)dst, (const Tin*)src, alpha_log, beta_log, tp, B, T, C, ld, L, (int)n_int, kEdge, nitems, taps, (int)main_blocks,
          nseg, nseg_head, tail_start, nitems_edge);
    else
      launch_pdl(act1d_cl_packed_kernel<Tin, Tout, FAST, false>, (unsigned)blocks, kPackedThreads, 0, st,
          (Tout*)dst, (const Tin*)src, alpha_log, beta_log, tp, B, T, C, ld, L, (int)n_int, kEdge, nitems, taps, (int)main_blocks,
          nseg, nseg_head, tail_start, nitems_edge);
    BVG_LAUNCHED();
    return BVG_OK;
  }
  if (ld > C) BVG_FAIL(BVG_EINVAL, "act1d_cl: a row pitch wider than the channel count needs T >= 49");
  const int64_t blocks = ceil_div(nitems_edge, threads);
  if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d_cl: tensor too large (%lld blocks)", (long long)blocks);
  launch_pdl(act1d_cl_kernel<Tin, Tout, 2, FAST>, (unsigned)blocks, threads, 0, st,
      (Tout*)dst, (const Tin*)src, alpha_log, beta_log, taps, B, T, C, ld, kEdge, nseg, nseg_head, T, nitems_edge);
  BVG_LAUNCHED();
  return BVG_OK;
}

// host entry used by the C ABI and by the vocoder plan
int act1d_cl_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                    const Taps& taps, int B, int64_t T, int C, int in_dtype, int out_dtype, bool fast,
                    cudaStream_t st, int ld) {
  if (B <= 0 || T <= 0 || C <= 0) return BVG_OK;
  if (ld <= 0) ld = C;
  if (ld < C) BVG_FAIL(BVG_EINVAL, "act1d_cl: row pitch %d < channels %d", ld, C);
  // ld > C: exactly 8 pad channels per row are supported (zero-filled, store_pad8); rows must then be 16-byte aligned
  const int out_es = out_dtype == BVG_BF16 ? 2 : 4;
  if (ld > C && (ld != C + 8 || (C & 1) || (ld * out_es) % 16 || (C * out_es) % 16 || reinterpret_cast<uintptr_t>(dst) % 16))
    BVG_FAIL(BVG_EINVAL, "act1d_cl: unsupported row pitch %d for %d channels", ld, C);
  typedef __nv_bfloat16 bf;
  if (in_dtype == BVG_F32 && out_dtype == BVG_F32)
    return fast ? launch_cl<float, float, true>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st)
                : launch_cl<float, float, false>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st);
  if (in_dtype == BVG_F32 && out_dtype == BVG_BF16)
    return fast ? launch_cl<float, bf, true>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st)
                : launch_cl<float, bf, false>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st);
  if (in_dtype == BVG_BF16 && out_dtype == BVG_BF16)
    return fast ? launch_cl<bf, bf, true>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st)
                : launch_cl<bf, bf, false>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st);
  if (in_dtype == BVG_BF16 && out_dtype == BVG_F32)
    return fast ? launch_cl<bf, float, true>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st)
                : launch_cl<bf, float, false>(dst, src, alpha_log, beta_log, taps, B, T, C, ld, st);
  BVG_FAIL(BVG_EDTYPE, "act1d_cl: unsupported dtype pair (%d -> %d)", in_dtype, out_dtype);
}

}  // namespace bvg
