// Fused anti-aliased activation on the reference layout [B, C, T] (time fastest).
// Replaces the reference's `anti_alias_activation_cuda.forward`
// (alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-246) with the exact
// edge semantics of the torch operator (alias_free_activation/torch/act.py:25-30).
//
// One CTA = one time tile of one (b,c) row:
//   1. the tile plus an 8-sample halo is staged in shared memory by ONE bulk
//      async copy (cp.async.bulk / UBLKCP, mbarrier completion) when the row
//      pitch is 16-byte aligned, else by cooperative scalar loads;
//   2. the tile is cut into 256 odd-length segments (odd stride => conflict-free
//      LDS/STS); thread i walks segments i and i+128 TOGETHER as one f32x2 pair
//      (FFMA2: two FMAs per issue slot - the kernel is FP32-issue bound, not HBM
//      bound, see profiles/) with the rotating 6+12 register window: 24 FMA + 2
//      snake per output, no recomputation inside a segment.  Segments that touch
//      a row end (replicate clamps, v edge rules) or are short take the scalar
//      generic routine;
//   3. outputs are staged in shared memory and leave with ONE bulk async store
//      (or cooperative stores on the unaligned path).
#include "act_packed.cuh"

namespace bvg {

constexpr int kBctThreads = 128;
constexpr int kBctSegs = 2 * kBctThreads;   // segments per tile
constexpr int kBctMaxSeg = 31;              // 6n-5
constexpr int kBctHalo = 8;                 // >= 5, multiple of 8 elements => 16 B for bf16, 32 B for fp32

// generic scalar segment [t0, t1): any position, any length
template <typename T, bool FAST>
__device__ __forceinline__ void bct_segment_scalar(const T* s_in, T* s_out, const Taps& taps, float a, float ib,
                                                   int64_t t0, int64_t t1, int64_t tile_t0, int64_t lo, int n_in,
                                                   int64_t Tlen) {
  const int64_t tlast = Tlen - 1;
  float X[6], V[12], vend = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    int64_t ti = t0 - 5 + i;
    ti = ti < 0 ? 0 : (ti > tlast ? tlast : ti);
    X[i] = to_f32<T>(s_in[ti - lo]);
  }
  const int nsteps = (int)(t1 - t0) + 5;
  for (int base = 0; base < nsteps; base += 6) {
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int64_t t = t0 - 5 + base + s;
      int64_t tl = t + 5;
      tl = tl > tlast ? tlast : tl;
      // past the end of the segment the index may leave the staged range; those steps
      // produce nothing, so any in-range sample will do
      int idx = (int)(tl - lo);
      idx = idx < n_in ? idx : n_in - 1;
      X[(s + 5) % 6] = to_f32<T>(s_in[idx]);
      float uo = snake_acc_init<FAST>(ib), ue = uo;   // same operation order as the packed path
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float xv = X[(s + 5 - q) % 6];
        uo = fmaf(taps.up[2 * q], xv, uo);
        ue = fmaf(taps.up[2 * q + 1], xv, ue);
      }
      float vo = snake_apply<FAST>(uo, a, ib);
      float ve = snake_apply<FAST>(ue, a, ib);
      if (t >= Tlen - 3) {          // right edge: v[m >= 2T] := v[2T-1] (odd sample of step T-3)
        if (t == Tlen - 3) vend = vo;
        vo = vend;
        ve = vend;
      }
      V[(2 * s + 10) % 12] = vo;
      V[(2 * s + 11) % 12] = ve;
      if (s == 2 && base == 0 && t0 == 0) {   // left edge: v[m < 0] := v[0]
        const float v0 = V[3];
        V[10] = v0; V[11] = v0; V[0] = v0; V[1] = v0; V[2] = v0;
      }
      float acc = taps.down[0] * V[(2 * s) % 12];
#pragma unroll
      for (int k = 1; k < 12; ++k) acc = fmaf(taps.down[k], V[(2 * s + k) % 12], acc);
      if (t >= t0 && t < t1) s_out[t - tile_t0] = from_f32<T>(acc);
    }
  }
}

#define BVG_BCT2_STEP(S, WITH_DOWN, OIDX)                                                   \
  {                                                                                         \
    X[((S) + 5) % 6] = pk2(to_f32<T>(ipa[(S)]), to_f32<T>(ipb[(S)]));                       \
    f32x2 uo = sn.acc_init(), ue = uo;                                                      \
    _Pragma("unroll") for (int q = 0; q < 6; ++q) {                                         \
      const f32x2 xv = X[((S) + 5 - q) % 6];                                                \
      uo = fma2(tp.u[q], xv, uo);                                                           \
      ue = fma2(tp.u[5 - q], xv, ue);                                                       \
    }                                                                                       \
    V[(2 * (S) + 10) % 12] = sn.apply(uo);                                                  \
    V[(2 * (S) + 11) % 12] = sn.apply(ue);                                                  \
    if (WITH_DOWN) {                                                                        \
      f32x2 acc = mul2(tp.d[0], V[(2 * (S)) % 12]);                                         \
      _Pragma("unroll") for (int k = 1; k < 12; ++k)                                        \
        acc = fma2(tp.d[k < 6 ? k : 11 - k], V[(2 * (S) + k) % 12], acc);                   \
      float ya, yb;                                                                         \
      upk2(acc, ya, yb);                                                                    \
      opa[(OIDX)] = from_f32<T>(ya);                                                        \
      opb[(OIDX)] = from_f32<T>(yb);                                                        \
    }                                                                                       \
  }

template <typename T, bool FAST>
__global__ void __launch_bounds__(kBctThreads)
act1d_bct_kernel(T* __restrict__ dst, const T* __restrict__ src, const float* __restrict__ alpha_log,
                 const float* __restrict__ beta_log, const Taps taps, const TapsPacked tp, int C, int64_t Tlen, int L,
                 int tile_len, int tiles_per_row, int aligned) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int64_t row = blockIdx.x / tiles_per_row;
  const int tile = blockIdx.x % tiles_per_row;
  const int64_t tile_t0 = (int64_t)tile * tile_len;
  const int c = (int)(row % C);

  // staged input range [lo, hi) of this row
  const int64_t lo = tile_t0 - kBctHalo > 0 ? tile_t0 - kBctHalo : 0;
  int64_t hi = tile_t0 + tile_len + kBctHalo;
  if (hi > Tlen) hi = Tlen;
  const int n_in = (int)(hi - lo);
  T* s_in = reinterpret_cast<T*>(smem_raw);
  // output staging starts after the input region (rounded to 128 B)
  const int in_bytes = ((tile_len + 2 * kBctHalo) * (int)sizeof(T) + 127) & ~127;
  T* s_out = reinterpret_cast<T*>(smem_raw + in_bytes);

  const T* rsrc = src + row * Tlen;
  T* rdst = dst + row * Tlen;

  if (aligned) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      const uint32_t bytes = (uint32_t)n_in * sizeof(T);
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(s_in, rsrc + lo, bytes, &bar);
    }
  } else {
    for (int i = threadIdx.x; i < n_in; i += kBctThreads) s_in[i] = rsrc[lo + i];
  }

  const float al = __ldg(alpha_log + c), be = __ldg(beta_log + c);

  __syncthreads();  // makes the mbarrier init (or the cooperative loads) visible
  if (aligned) mbar_wait(&bar, 0);

  const int64_t tile_end = tile_t0 + tile_len < Tlen ? tile_t0 + tile_len : Tlen;
  const int64_t tlast = Tlen - 1;
  const int64_t t0a = tile_t0 + (int64_t)threadIdx.x * L;
  const int64_t t0b = t0a + (int64_t)kBctThreads * L;
  const bool fast_a = t0a >= 5 && t0a + L + 4 <= tlast && t0a + L <= tile_end;
  const bool fast_b = t0b >= 5 && t0b + L + 4 <= tlast && t0b + L <= tile_end;

  if (fast_a && fast_b) {
    // ---- both segments interior and full (L = 6n-5): packed pair, no clamps/predicates ----
    SnakePair<FAST> sn;
    sn.init(al, al, be, be);
    const T* ipa = s_in + (t0a - 5 - lo);
    const T* ipb = s_in + (t0b - 5 - lo);
    T* opa = s_out + (t0a - tile_t0);
    T* opb = s_out + (t0b - tile_t0);
    f32x2 X[6], V[12];
#pragma unroll
    for (int i = 0; i < 5; ++i) X[i] = pk2(to_f32<T>(ipa[i]), to_f32<T>(ipb[i]));
    ipa += 5;
    ipb += 5;
#pragma unroll
    for (int s = 0; s < 6; ++s) {   // first body: 5 warm-up steps + 1 full step
      if (s < 5) BVG_BCT2_STEP(s, false, 0) else BVG_BCT2_STEP(s, true, 0)
    }
    ipa += 6; ipb += 6; opa += 1; opb += 1;
    const int nbody = (L + 5) / 6 - 1;
    for (int it = 0; it < nbody; ++it) {
#pragma unroll
      for (int s = 0; s < 6; ++s) BVG_BCT2_STEP(s, true, s)
      ipa += 6; ipb += 6; opa += 6; opb += 6;
    }
  } else {
    const float a = expf(al);
    const float ib = 1.0f / (expf(be) + 1e-9f);
    int64_t t1a = t0a + L < tile_end ? t0a + L : tile_end;
    int64_t t1b = t0b + L < tile_end ? t0b + L : tile_end;
    if (t0a < t1a) bct_segment_scalar<T, FAST>(s_in, s_out, taps, a, ib, t0a, t1a, tile_t0, lo, n_in, Tlen);
    if (t0b < t1b) bct_segment_scalar<T, FAST>(s_in, s_out, taps, a, ib, t0b, t1b, tile_t0, lo, n_in, Tlen);
  }

  const int n_out = (int)(tile_end - tile_t0);
  if (aligned) {
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0 && n_out > 0) {
      // n_out*sizeof(T) is a multiple of 16: tile_len is a multiple of 256 and Tlen*sizeof(T) % 16 == 0
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rdst + tile_t0),
                   "r"(smem_u32(s_out)), "r"((uint32_t)n_out * (uint32_t)sizeof(T))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the read
    }
  } else {
    __syncthreads();
    for (int i = threadIdx.x; i < n_out; i += kBctThreads) rdst[tile_t0 + i] = s_out[i];
  }
}

template <typename T, bool FAST>
static int launch_bct(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                      const Taps& taps, int B, int C, int64_t Tlen, cudaStream_t st) {
  const int64_t rows = (int64_t)B * C;
  // tiles: as few as possible per row, then the shortest 6n-5 segment that covers them
  const int max_tile = kBctSegs * kBctMaxSeg;
  const int tiles_per_row = (int)ceil_div(Tlen, max_tile);
  const int64_t per_tile = ceil_div(Tlen, tiles_per_row);
  int L = (int)ceil_div(per_tile, kBctSegs);
  L = (int)ceil_div(L + 5, 6) * 6 - 5;  // 6n-5: whole 6-step bodies; odd => conflict-free shared-memory walk
  if (L < 7) L = 7;                 // only the first segment of a row may see v[m<0] (needs 2*L-5 >= 0)
  if (L > kBctMaxSeg) L = kBctMaxSeg;
  const int tile_len = kBctSegs * L;   // multiple of 256 elements
  const int64_t blocks = rows * tiles_per_row;
  if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d: tensor too large (%lld blocks)", (long long)blocks);
  const int aligned = ((Tlen * (int64_t)sizeof(T)) % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const int in_bytes = ((tile_len + 2 * kBctHalo) * (int)sizeof(T) + 127) & ~127;
  const int smem = in_bytes + tile_len * (int)sizeof(T);
  auto kern = act1d_bct_kernel<T, FAST>;
  if (smem > 48 * 1024)  // per device/context attribute; cheap enough to set on every large launch
    BVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  TapsPacked tp;
  make_taps_packed(&tp, taps);
  kern<<<(unsigned)blocks, kBctThreads, smem, st>>>((T*)dst, (const T*)src, alpha_log, beta_log, taps, tp, C, Tlen,
                                                    L, tile_len, tiles_per_row, aligned);
  BVG_LAUNCHED();
  return BVG_OK;
}

int act1d_bct_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                     const Taps& taps, int B, int C, int64_t T, int dtype, bool fast, cudaStream_t st) {
  if (B <= 0 || C <= 0 || T <= 0) return BVG_OK;
  if (dtype == BVG_F32)
    return fast ? launch_bct<float, true>(dst, src, alpha_log, beta_log, taps, B, C, T, st)
                : launch_bct<float, false>(dst, src, alpha_log, beta_log, taps, B, C, T, st);
  if (dtype == BVG_BF16)
    return fast ? launch_bct<__nv_bfloat16, true>(dst, src, alpha_log, beta_log, taps, B, C, T, st)
                : launch_bct<__nv_bfloat16, false>(dst, src, alpha_log, beta_log, taps, B, C, T, st);
  BVG_FAIL(BVG_EDTYPE, "act1d: unsupported dtype %d", dtype);
}

}  // namespace bvg
