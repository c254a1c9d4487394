// Fused anti-aliased activation on the reference layout [B, C, T] (time fastest).
// Replaces the reference's `anti_alias_activation_cuda.forward`
// (alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-246) with the exact
// edge semantics of the torch operator (alias_free_activation/torch/act.py:25-30).
//
// One CTA = one time tile of one (b,c) row:
//   1. the tile plus an 8-sample halo is staged in shared memory by ONE bulk
//      async copy (cp.async.bulk / UBLKCP, mbarrier completion) when the row
//      pitch is 16-byte aligned, else by cooperative scalar loads;
//   2. the tile is cut into <= 256 odd-length segments (odd stride => conflict-free
//      LDS/STS); a thread walks two segments TOGETHER as one f32x2 pair
//      (FFMA2: two FMAs per issue slot - the kernel is FP32-issue bound, not HBM
//      bound, see profiles/) with the rotating 6+12 register window: 24 FMA + 2
//      snake per output, no recomputation inside a segment.  The few outputs next
//      to a row end (replicate clamps, v edge rules) are evaluated one per thread
//      straight from the definition (bit-identical operation order);
//   3. outputs are staged in shared memory and leave with ONE bulk async store
//      (or cooperative stores on the unaligned path).
#include "act_packed.cuh"

namespace bvg {

constexpr int kBctThreads = 128;
constexpr int kBctSegs = 2 * kBctThreads;   // segments per tile
constexpr int kBctMaxSeg = 31;              // 6n-5
constexpr int kBctHalo = 8;                 // >= 5, multiple of 8 elements => 16 B for bf16, 32 B for fp32

// One output straight from the definition (clamped indices = the replicate rules of the torch operator), with the operation
// order of the sliding-window routines: v[2tau+5] and v[2tau+6] are the "odd"/"even" results of window step tau
//   v[2tau+5] = snake(init + sum_{q=0..5} up[2q]   * x[clamp(tau+5-q)])      (fma chain, q ascending)
//   v[2tau+6] = snake(init + sum_{q=0..5} up[2q+1] * x[clamp(tau+5-q)])
//   y[t]      = down[0]*v[c(2t-5)] then fma(down[k], v[c(2t-5+k)], .) for k = 1..11,   c(m) = clamp(m, 0, 2T-1)
template <typename T, bool FAST>
__device__ __forceinline__ float bct_output_direct(const T* s_in, const Taps& taps, float a, float ib, int64_t t, int64_t lo,
                                                   int64_t Tlen) {
  const int64_t tlast = Tlen - 1, mlast = 2 * Tlen - 1;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    int64_t m = 2 * t - 5 + k;
    m = m < 0 ? 0 : (m > mlast ? mlast : m);
    const int odd = (int)(m & 1);
    const int64_t tau = odd ? (m - 5) / 2 : (m - 6) / 2;   // exact: m - 5 / m - 6 are even
    float u = snake_acc_init<FAST>(ib);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      int64_t xi = tau + 5 - q;
      xi = xi < 0 ? 0 : (xi > tlast ? tlast : xi);
      const float xv = to_f32<T>(s_in[xi - lo]);
      u = fmaf(odd ? taps.up[2 * q] : taps.up[2 * q + 1], xv, u);
    }
    const float v = snake_apply<FAST>(u, a, ib);
    acc = k == 0 ? taps.down[0] * v : fmaf(taps.down[k], v, acc);
  }
  return acc;
}

#define BVG_BCT2_STEP(S, WITH_DOWN, OIDX) BVG_BCT2_STEP_X(S, WITH_DOWN, OIDX, pk2(to_f32<T>(ipa[(S)]), to_f32<T>(ipb[(S)])))
#define BVG_BCT2_STEP_X(S, WITH_DOWN, OIDX, XNEW)                                           \
  {                                                                                         \
    X[((S) + 5) % 6] = (XNEW);                                                              \
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
  // Work split of a tile.  FAST segments (full length L, no index clamp anywhere in their 5-step warm-up and run-out) are
  // segments [s_first, s_tail) and are walked in pairs by the packed routine: thread i < H = ceil(nfast / 2) takes
  // segments s_first + i and s_first + i + H (pairing at distance H, not at a fixed 128, keeps both halves of every pair
  // full in a partly filled tile; an odd segment out is walked twice by its thread, which costs nothing extra because
  // the rest of its warp is in the same loop).  SLOW outputs - the first segment of a row and the last one or two,
  // where the replicate rules of the torch operator apply - are NOT walked: each of them (at most ~3 L per tile) is
  // evaluated on its own by one thread, starting from the top of the block where threads are idle, straight from the
  // definition with clamped indices and the SAME operation order as the sliding window, so the value is bit-identical.
  // (Before: the scalar edge-aware walk of two segments cost as many warp instructions as the whole rest of the tile -
  //  ncu, [43,192,8192] fp32: 154 M warp instructions for 68 M elements against 73 M for 76 M elements at T = 131072 -
  //  because the warp holding an edge thread runs both paths one after the other, and with T = 8192 every tile has an edge.)
  const int s_first = tile_t0 == 0 ? 1 : 0;
  const int len = (int)(tile_end - tile_t0);         // 32-bit from here on: a 64-bit division costs ~100 instructions per thread
  int s_tail = len / L;
  {
    const int64_t room64 = tlast - 4 - tile_t0;      // segments must also end 4 samples before the end of the row
    const int room = room64 > (int64_t)len ? len : (room64 < 0 ? 0 : (int)room64);
    const int s2 = room / L;
    if (s2 < s_tail) s_tail = s2;
  }
  const int nfast = s_tail > s_first ? s_tail - s_first : 0;
  const int H = (nfast + 1) >> 1;

  if ((int)threadIdx.x < H) {
    const int sa = s_first + (int)threadIdx.x;
    int sb = sa + H;
    if (sb >= s_tail) sb = sa;
    const int64_t t0a = tile_t0 + (int64_t)sa * L, t0b = tile_t0 + (int64_t)sb * L;
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
  }
  {
    const int64_t head_end = tile_t0 == 0 ? (L < tile_end ? L : tile_end) : tile_t0;
    int64_t tail_start = tile_t0 + (int64_t)(s_tail > 0 ? s_tail : 0) * L;
    if (tail_start < head_end) tail_start = head_end;
    const int nhead = (int)(head_end - tile_t0);
    const int nslow = nhead + (int)(tile_end - tail_start);
    if (nslow > 0) {
      const float a = expf(al);
      const float ib = 1.0f / (expf(be) + 1e-9f);
      for (int j = kBctThreads - 1 - (int)threadIdx.x; j < nslow; j += kBctThreads) {
        const int64_t t = j < nhead ? tile_t0 + j : tail_start + (j - nhead);
        s_out[t - tile_t0] = from_f32<T>(bct_output_direct<T, FAST>(s_in, taps, a, ib, t, lo, Tlen));
      }
    }
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

// ---------------------------------------------------------------------------------------------------------------------
// In-place variant (round 2): outputs overwrite the staged input (y[t] goes where x[t] was), so a CTA needs ONE tile buffer
// instead of two.  For fp32 tiles of ~7 900 samples that is 32 KB instead of 64 KB: 6 CTAs (24 warps) per SM instead of 3,
// +25-34 % on rows of >= 131 072 samples (same box, `tools/act_bct_points.py`: 0.47 -> 0.59 of the HBM copy rate at T = 131 072,
// 0.68 -> 0.91 at T = 2 M, 192 channels).  A segment's own reads run 5 samples ahead of its writes; what it needs from its
// neighbours' segments - the 5 warm-up samples on its left, the 5 on its right - and every edge output are read BEFORE a
// block-wide barrier, the first write comes after it.  Short tiles (T = 8 192: every tile holds a row end, and two buffers
// already fit 6 CTAs) are 8-12 % slower this way and keep the two-buffer kernel above; so does 16-bit I/O.
template <typename T, bool FAST>
__global__ void __launch_bounds__(kBctThreads, 6)
act1d_bct_inplace_kernel(T* __restrict__ dst, const T* __restrict__ src, const float* __restrict__ alpha_log,
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
  // outputs overwrite the inputs in place: y[t] at the position of x[t] (tile_t0 - lo is 0 or kBctHalo elements: 16-byte aligned)
  T* s_out = s_in + (tile_t0 - lo);

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
  // Work split of a tile.  FAST segments (full length L, no index clamp anywhere in their 5-step warm-up and run-out) are
  // segments [s_first, s_tail) and are walked in pairs by the packed routine: thread i < H = ceil(nfast / 2) takes
  // segments s_first + i and s_first + i + H (pairing at distance H, not at a fixed 128, keeps both halves of every pair
  // full in a partly filled tile; an odd segment out is walked twice by its thread, which costs nothing extra because
  // the rest of its warp is in the same loop).  SLOW outputs - the first segment of a row and the last one or two,
  // where the replicate rules of the torch operator apply - are NOT walked: each of them (at most ~3 L per tile) is
  // evaluated on its own by one thread, starting from the top of the block where threads are idle, straight from the
  // definition with clamped indices and the SAME operation order as the sliding window, so the value is bit-identical.
  // (Before: the scalar edge-aware walk of two segments cost as many warp instructions as the whole rest of the tile -
  //  ncu, [43,192,8192] fp32: 154 M warp instructions for 68 M elements against 73 M for 76 M elements at T = 131072 -
  //  because the warp holding an edge thread runs both paths one after the other, and with T = 8192 every tile has an edge.)
  const int s_first = tile_t0 == 0 ? 1 : 0;
  const int len = (int)(tile_end - tile_t0);         // 32-bit from here on: a 64-bit division costs ~100 instructions per thread
  int s_tail = len / L;
  {
    const int64_t room64 = tlast - 4 - tile_t0;      // segments must also end 4 samples before the end of the row
    const int room = room64 > (int64_t)len ? len : (room64 < 0 ? 0 : (int)room64);
    const int s2 = room / L;
    if (s2 < s_tail) s_tail = s2;
  }
  const int nfast = s_tail > s_first ? s_tail - s_first : 0;
  const int H = (nfast + 1) >> 1;

  // ---- phase A: everything that reads samples another thread will overwrite --------------------------------------------
  // (i) the edge output of this thread, if any: the first segment of a row and the last one or two, < 3 L + 8 outputs per
  //     tile, i.e. at most one per thread (from the top of the block, where threads have no segments to walk)
  static_assert(3 * kBctMaxSeg + 8 <= kBctThreads, "one edge output per thread");
  const int64_t head_end = tile_t0 == 0 ? (L < tile_end ? L : tile_end) : tile_t0;
  int64_t tail_start = tile_t0 + (int64_t)(s_tail > 0 ? s_tail : 0) * L;
  if (tail_start < head_end) tail_start = head_end;
  const int nhead = (int)(head_end - tile_t0);
  const int nslow = nhead + (int)(tile_end - tail_start);
  const int jslow = kBctThreads - 1 - (int)threadIdx.x;
  float slow_v = 0.f;
  int64_t slow_t = -1;
  if (jslow < nslow) {
    const float a = expf(al);
    const float ib = 1.0f / (expf(be) + 1e-9f);
    slow_t = jslow < nhead ? tile_t0 + jslow : tail_start + (jslow - nhead);
    slow_v = bct_output_direct<T, FAST>(s_in, taps, a, ib, slow_t, lo, Tlen);
  }
  // (ii) the window warm-up (5 samples left of the segment) and the 5 samples right of it, both owned by other segments
  const bool walker = (int)threadIdx.x < H;
  const int sa = s_first + (int)threadIdx.x;
  int sb = sa + H;
  if (sb >= s_tail) sb = sa;
  const int64_t t0a = tile_t0 + (int64_t)sa * L, t0b = tile_t0 + (int64_t)sb * L;
  const T* ipa = s_in + (t0a - 5 - lo);
  const T* ipb = s_in + (t0b - 5 - lo);
  f32x2 X[6], V[12], RH[5];
  if (walker) {
#pragma unroll
    for (int i = 0; i < 5; ++i) X[i] = pk2(to_f32<T>(ipa[i]), to_f32<T>(ipb[i]));
#pragma unroll
    for (int i = 0; i < 5; ++i) RH[i] = pk2(to_f32<T>(ipa[5 + L + i]), to_f32<T>(ipb[5 + L + i]));
  }
  __syncthreads();
  // ---- phase B: writes (nobody reads a sample of another segment from here on)
  if (slow_t >= 0) s_out[slow_t - tile_t0] = from_f32<T>(slow_v);
  if (walker) {
    SnakePair<FAST> sn;
    sn.init(al, al, be, be);
    T* opa = s_out + (t0a - tile_t0);
    T* opb = s_out + (t0b - tile_t0);
    ipa += 5;
    ipb += 5;
#pragma unroll
    for (int s = 0; s < 6; ++s) {   // first body: 5 warm-up steps + 1 full step
      if (s < 5) BVG_BCT2_STEP(s, false, 0) else BVG_BCT2_STEP(s, true, 0)
    }
    ipa += 6; ipb += 6; opa += 1; opb += 1;
    const int nbody = (L + 5) / 6 - 1;   // >= 1 (L >= 7)
    for (int it = 0; it < nbody - 1; ++it) {
#pragma unroll
      for (int s = 0; s < 6; ++s) BVG_BCT2_STEP(s, true, s)
      ipa += 6; ipb += 6; opa += 6; opb += 6;
    }
    // last body: its first load is the segment's own last sample, the other five come from the registers read in phase A
    BVG_BCT2_STEP(0, true, 0)
#pragma unroll
    for (int s = 1; s < 6; ++s) BVG_BCT2_STEP_X(s, true, s, RH[s - 1])
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
  // equal tiles (a multiple of 256 elements keeps every tile start 16-byte aligned for the bulk copies), cut into the
  // shortest 6n-5 segments that cover one tile with at most 256 of them
  const int64_t per_tile = ceil_div(ceil_div(Tlen, tiles_per_row), 256) * 256;
  int L = (int)ceil_div(per_tile, kBctSegs);
  L = (int)ceil_div(L + 5, 6) * 6 - 5;  // 6n-5: whole 6-step bodies; odd => conflict-free shared-memory walk
  if (L < 7) L = 7;                 // only the first segment of a row may see v[m<0] (needs 2*L-5 >= 0)
  if (L > kBctMaxSeg) L = kBctMaxSeg;
  const int tile_len = (int)(per_tile < (int64_t)kBctSegs * L ? per_tile : (int64_t)kBctSegs * L);
  const int64_t blocks = rows * tiles_per_row;
  if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d: tensor too large (%lld blocks)", (long long)blocks);
  const int aligned = ((Tlen * (int64_t)sizeof(T)) % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const int in_bytes = ((tile_len + 2 * kBctHalo) * (int)sizeof(T) + 127) & ~127;
  // one buffer (in place) when two would limit the SM to fewer than 6 CTAs, i.e. for long fp32 tiles
  const bool inplace = sizeof(T) == 4 && in_bytes + tile_len * (int)sizeof(T) > 36 * 1024;
  const int smem = inplace ? in_bytes : in_bytes + tile_len * (int)sizeof(T);
  auto kern = inplace ? act1d_bct_inplace_kernel<T, FAST> : act1d_bct_kernel<T, FAST>;
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
  if (dtype == BVG_F16)   // the reference kernel dispatches half as well (type_shim.h:20-43); the arithmetic is fp32 as for bf16
    return fast ? launch_bct<__half, true>(dst, src, alpha_log, beta_log, taps, B, C, T, st)
                : launch_bct<__half, false>(dst, src, alpha_log, beta_log, taps, B, C, T, st);
  BVG_FAIL(BVG_EDTYPE, "act1d: unsupported dtype %d", dtype);
}

}  // namespace bvg
