// One AMPBlock1 unit (bigvgan.py:132-141, one iteration of the loop) as ONE kernel for the narrow stages
// (<= 96 channels):
//
//   y = x + c2( a2( c1( a1(x) ) ) )            a1 / a2 = anti-aliased SnakeBeta (Activation1d), c1 dilated, c2 dilation 1
//
// x and y are the fp32 residual stream in channels-last HBM [B, T, Cp]; everything between them lives on the SM:
//
//   phase A1   x (global, fp32) -> up x2 / snake / down x2 in registers (sliding window along time, one channel
//              PAIR per thread as f32x2, the scheme of act1d_cl.cu) -> bf16 written straight into the K-major
//              SWIZZLE_128B operand tile of conv1 (thread-written swizzle, tools/swz_probe.cu)
//   MMA 1      tcgen05.mma, M = 128 (channels, narrow layers replicated 2x / 4x as in conv_umma2.cu), N = N1 <= 256
//              time columns, all k taps through row-shifted descriptors of the one tile; weights stream from L2
//              through a TMA ring; fp32 accumulator in TMEM
//   phase A2   TMEM (lane = channel, column = time) -> + bias -> Activation1d along the registers of a thread (the
//              epilogue scheme of conv_umma2a.cu: two lockstep sub-segments per thread packed in f32x2) -> bf16
//              written into the SAME shared-memory tile (conv1 has finished reading it) as the operand of conv2
//   MMA 2      conv2 over that tile, accumulator in the same TMEM columns
//   phase ST   TMEM + bias + residual x (global; an L2 hit, the tile was read in phase A1) -> y
//              (optionally (.)*scale + accum and/or a bf16 result: the resblock mean of bigvgan.py:369-375 and
//              the next stage's ConvTranspose input, as in the conv_umma2 epilogue)
//
// HBM traffic per element and unit: 4 B read (+ halo) + 4 B written, against 20 B for the layer-by-layer form.
// A CTA runs its phases one after the other; TWO CTAs are resident per SM (<= 113 KB shared memory, 256 TMEM
// columns, <= 102 registers each), so the FP32 phases of one overlap the tensor-core phases of the other.
//
// Tile geometry (host plan below): the a2 rows of a tile are split into 4*rep sub-segments of L2 rows (bodies of 6
// steps), R2T = 4*rep*L2; conv1 computes N1 >= R2T + 11 columns (5 + 5 halo of a2, + 1 lead step), conv2 emits
// NOUT = floor16(R2T - (k-1)) outputs, a1 produces R1 = N1 + (k-1)*dil rows.
//
// Sequence ends follow the torch operator (replicate padding of x and of the activated upsampled signal, zero padding of
// both convolutions): tiles that touch an end run the same sliding windows with clamped loads, write zero rows outside
// [0, T) and then recompute the 3 + 3 outputs next to the ends straight from the definition.
#include "act_packed.cuh"
#include "umma_common.cuh"

namespace bvg {

constexpr int AU_CWARPS = 8;                          // compute warps (0..7); warp 8: weight loads, warp 9: MMA issue
constexpr int AU_CTHREADS = 32 * AU_CWARPS;
constexpr int AU_THREADS = AU_CTHREADS + 64;
constexpr int AU_MAX_SMEM = 113 * 1024;               // two CTAs per SM: 2 x (113 KB + 1 KB system) = 228 KB
constexpr int AU_MAX_SLOTS = 6;
constexpr int AU_TMEM_COLS = 256;

struct AUParams {
  const float* x;            // [B, T, ld] fp32
  void* out;                 // [B, T, ld] fp32 (bf16 if out_bf16)
  const float* accum;        // optional fp32 [B, T, ld]
  float scale;
  int out_bf16;
  const float *bias1, *bias2;                 // [128] fp32, zero in pad rows
  const float *al1, *be1, *al2, *be2;         // [Cp] log-scale snake parameters (pad entries 0)
  TapsPacked tp1, tp2;
  Taps t1, t2;
  int B, T, C, Cp, ld;
  int k, dil, h1, h2;
  int nch;                   // 64-channel K chunks
  int rep, LR, wrows;
  int P, zpad, NSEG1, L1;    // phase A1: channel pairs, pad-chunk flag, segments, rows per segment
  int L2, R2T;               // phase A2: rows per sub-segment, rows per tile
  int N1, N2, NOUT, R1, RB;
  int chb;                   // bytes per 64-channel chunk of the operand tile
  int nslot, slotb;          // weight ring
  int ncol_st;               // phase ST: columns per warp
  int n_ttiles;
  int64_t n_tiles;
};

struct SnakeC {
  f32x2 hb, nhb, a2, na2hb;
  __device__ __forceinline__ void init(float a0, float a1, float ib0, float ib1) {
    hb = pk2(0.5f * ib0, 0.5f * ib1);
    nhb = pk2(-0.5f * ib0, -0.5f * ib1);
    a2 = pk2(2.0f * a0, 2.0f * a1);
    na2hb = pk2(-2.0f * a0 * 0.5f * ib0, -2.0f * a1 * 0.5f * ib1);
  }
};

__device__ __forceinline__ f32x2 au_add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t bf16x2_bits(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One step of the sliding window (time t): XIN = x[t+5] enters the window; v[2t+5], v[2t+6] are produced.
// X / V / sn / tp are the caller's locals; S is the step index modulo 6 (a constant after unrolling).
#define AU_STEP_UP(S, XIN)                                                        \
  {                                                                                \
    X[((S) + 5) % 6] = (XIN);                                                      \
    f32x2 uo_ = sn.hb, ue_ = sn.hb;                                                \
    _Pragma("unroll") for (int q_ = 0; q_ < 6; ++q_) {                             \
      const f32x2 xv_ = X[((S) + 5 - q_) % 6];                                     \
      uo_ = fma2(tp.u[q_], xv_, uo_);                                              \
      ue_ = fma2(tp.u[5 - q_], xv_, ue_);                                          \
    }                                                                              \
    float zo0_, zo1_, ze0_, ze1_;                                                  \
    upk2(fma2(sn.a2, uo_, sn.na2hb), zo0_, zo1_);                                  \
    upk2(fma2(sn.a2, ue_, sn.na2hb), ze0_, ze1_);                                  \
    V[(2 * (S) + 10) % 12] = fma2(sn.nhb, pk2(__cosf(zo0_), __cosf(zo1_)), uo_);   \
    V[(2 * (S) + 11) % 12] = fma2(sn.nhb, pk2(__cosf(ze0_), __cosf(ze1_)), ue_);   \
  }
// y[t] of the same step: the 12-tap decimating FIR as two independent 6-term chains (even / odd taps)
#define AU_STEP_DOWN(S, YOUT)                                                      \
  {                                                                                \
    f32x2 ae_ = mul2(tp.d[0], V[(2 * (S)) % 12]);                                  \
    f32x2 ao_ = mul2(tp.d[1], V[(2 * (S) + 1) % 12]);                              \
    _Pragma("unroll") for (int k_ = 2; k_ < 12; k_ += 2) {                         \
      ae_ = fma2(tp.d[k_ < 6 ? k_ : 11 - k_], V[(2 * (S) + k_) % 12], ae_);        \
      ao_ = fma2(tp.d[k_ + 1 < 6 ? k_ + 1 : 10 - k_], V[(2 * (S) + k_ + 1) % 12], ao_); \
    }                                                                              \
    YOUT = au_add2(ae_, ao_);                                                      \
  }

// byte offset of (row, channel c) inside the K-major SWIZZLE_128B operand tile (64-channel chunks of `chb` bytes)
__device__ __forceinline__ uint32_t au_tile_off(int row, int c, int chb) {
  const int kk = c & 63;
  return (uint32_t)((c >> 6) * chb + row * 128 + ((((kk >> 3) ^ (row & 7))) << 4) + (kk & 7) * 2);
}

// One output of Activation1d straight from its definition (alias_free_activation/torch/{resample.py:29-38,55-58,
// filter.py:94-101, act.py:25-30}, activations.py:107-120), for the outputs next to a sequence end:
//   xw[i] = x[clamp(t - 5 + i, 0, T-1)], i = 0..10
//   vv[j] = v[2t - 5 + j] on the virtual (unclamped) axis; v[m < 0] := v[0], v[m > 2T-1] := v[2T-1]
__device__ __forceinline__ float au_act_point(const Taps& tp, float a, float ib, int t, int T, const float (&xw)[11]) {
  float vv[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float u = snake_acc_init<true>(ib);
    if ((j & 1) == 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) u = fmaf(tp.up[2 * q], xw[j / 2 + 5 - q], u);       // u[2tau+1], tau = t - 3 + j/2
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) u = fmaf(tp.up[2 * q + 1], xw[(j + 9) / 2 - q], u);  // u[2tau], tau = t + (j-5)/2
    }
    vv[j] = snake_apply<true>(u, a, ib);
  }
  const int i0 = 5 - 2 * t;              // index of v[0]      (clamp active for j < i0; only when t <= 2)
  const int i1 = 2 * (T - t) + 4;        // index of v[2T-1]   (clamp active for j > i1; only when t >= T-3)
  const float v0 = t == 0 ? vv[5] : (t == 1 ? vv[3] : vv[1]);
  const int e = T - 1 - t;
  const float v1 = e == 0 ? vv[6] : (e == 1 ? vv[8] : vv[10]);
  float y = 0.f;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float v = vv[j];
    if (j < i0) v = v0;
    if (j > i1) v = v1;
    y = fmaf(tp.down[j], v, y);
  }
  return y;
}

// ------------------------------------------------------------------------------------------------ phase A1
// thread (pair, seg): channel pair c0, c0+1; output rows [seg*L1, (seg+1)*L1) of the tile (time tA1 + row).
template <bool EDGE>
__device__ __forceinline__ void au_a1_phase(const AUParams& p, uint32_t tile_u32, int b, int tA1, int tid) {
  if (tid >= p.P * p.NSEG1) return;
  const TapsPacked& tp = p.tp1;
  const int seg = tid / p.P, pair = tid - seg * p.P;
  const int c0 = 2 * pair;
  SnakeC sn;
  sn.init(expf(__ldg(p.al1 + c0)), expf(__ldg(p.al1 + c0 + 1)), 1.0f / (expf(__ldg(p.be1 + c0)) + 1e-9f),
          1.0f / (expf(__ldg(p.be1 + c0 + 1)) + 1e-9f));
  const int T = p.T, Tlast = p.T - 1;
  const int row0 = seg * p.L1;
  const int tf = tA1 + row0;                           // time of this segment's first output
  const int64_t ld = p.ld;
  const float* xb = p.x + (int64_t)b * T * ld + c0;
  auto ldx = [&](int t) -> float2 {
    if (EDGE) t = t < 0 ? 0 : (t > Tlast ? Tlast : t);
    return BVG_LDG(reinterpret_cast<const float2*>(xb + (int64_t)t * ld));
  };
  f32x2 X[6], V[12];
  float2 R[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) { const float2 v = ldx(tf - 6 + i); X[i] = pk2(v.x, v.y); }
#pragma unroll
  for (int i = 0; i < 6; ++i) R[i] = ldx(tf - 1 + i);
  X[5] = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 12; ++i) V[i] = pk2(0.f, 0.f);
  const int kk = c0 & 63;
  const uint32_t obase = tile_u32 + (uint32_t)((c0 >> 6) * p.chb + (kk & 7) * 2);
  const uint32_t och = (uint32_t)(kk >> 3);
  const bool zpad = p.zpad && pair == p.P - 1;          // this thread also clears the 8 pad channels behind the last pair
  const uint32_t zbase = tile_u32 + (uint32_t)((p.C >> 6) * p.chb);
  const uint32_t zch = (uint32_t)((p.C & 63) >> 3);
  const int nbody = p.L1 / 6 + 1;
  const int RB = p.RB;
  for (int j = 0; j < nbody; ++j) {
    const bool more = j + 1 < nbody;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const f32x2 xin = pk2(R[s].x, R[s].y);
      if (more) R[s] = ldx(tf + 5 + 6 * j + s);
      AU_STEP_UP(s, xin)
      if (j >= 1) {
        f32x2 y;
        AU_STEP_DOWN(s, y)
        const int row = row0 + 6 * j + s - 6;
        if (row < RB) {
          float ya, yb;
          upk2(y, ya, yb);
          if (EDGE) {
            const int t = tA1 + row;
            if (t < 0 || t > Tlast) { ya = 0.f; yb = 0.f; }
          }
          const uint32_t rsw = (uint32_t)(row & 7);
          st_shared_b32(obase + (uint32_t)row * 128u + ((och ^ rsw) << 4), bf16x2_bits(ya, yb));
          if (zpad) st_shared_zero16(zbase + (uint32_t)row * 128u + ((zch ^ rsw) << 4));
        }
      }
    }
  }
}

// the <= 6 a1 outputs next to the sequence ends, one (channel, row) per thread
__device__ __forceinline__ void au_a1_patch(const AUParams& p, uint32_t tile_u32, int b, int tA1, int tid) {
  const int T = p.T, Tlast = p.T - 1;
  const int nc = 2 * p.P;
  const int rows = p.NSEG1 * p.L1 < p.RB ? p.NSEG1 * p.L1 : p.RB;
  for (int item = tid; item < nc * 6; item += AU_CTHREADS) {
    const int c = item % nc, q = item / nc;
    const int t = q < 3 ? q : T - 6 + q;
    if (t < 0 || t > Tlast) continue;
    if (q >= 3 && t < 3) continue;                      // already covered by the head rows (T < 6)
    const int row = t - tA1;
    if (row < 0 || row >= rows) continue;
    const float* xb = p.x + (int64_t)b * T * p.ld + c;
    float xw[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      int ti = t - 5 + i;
      ti = ti < 0 ? 0 : (ti > Tlast ? Tlast : ti);
      xw[i] = BVG_LDG(xb + (int64_t)ti * p.ld);
    }
    const float a = expf(__ldg(p.al1 + c)), ib = 1.0f / (expf(__ldg(p.be1 + c)) + 1e-9f);
    const float y = au_act_point(p.t1, a, ib, t, T, xw);
    st_shared_b16(tile_u32 + au_tile_off(row, c, p.chb), __bfloat16_as_ushort(__float2bfloat16_rn(y)));
  }
}

// ------------------------------------------------------------------------------------------------ phase A2
// warp (lane group g, half): lanes = 32 channels; rows [r0, r0 + 2*L2) of the tile as two lockstep sub-segments.
// Column n of the accumulator is time tA2 - 6 + n; a2 row r is time tA2 + r and needs columns r + 1 .. r + 11.
template <bool EDGE>
__device__ __forceinline__ void au_a2_phase(const AUParams& p, uint32_t tile_u32, uint32_t tmem_base, int tA2, int warp,
                                            int lane) {
  const TapsPacked& tp = p.tp2;
  const int g = warp & 3, half = warp >> 2;
  const int lane0 = (g * 32) % p.LR;
  if (lane0 >= p.Cp) return;
  const int replica = (g * 32) / p.LR;
  const int ch = lane0 + lane;
  const bool lane_ok = ch < p.Cp;
  const int L = p.L2, T = p.T, Tlast = p.T - 1;
  const int r0 = (replica * 2 + half) * 2 * L;
  const uint32_t lanebase = tmem_base + ((uint32_t)(g * 32) << 16);
  const uint32_t taddr = lanebase + (uint32_t)r0;
  const int tM0 = tA2 - 6;                              // time of accumulator column 0
  const int tsub = tA2 + r0 - 6;                        // time of column 0 of sub-segment A (B: + L)
  float bv = 0.f, a = 1.f, ib = 1.f;
  if (lane_ok) {
    bv = __ldg(p.bias1 + ch);
    a = expf(__ldg(p.al2 + ch));
    ib = 1.0f / (expf(__ldg(p.be2 + ch)) + 1e-9f);
  }
  SnakeC sn;
  sn.init(a, a, ib, ib);
  const f32x2 bv2 = pk2(bv, bv);
  // column fetches: c = column relative to the sub-segment's first column
  auto colclamp = [&](int t) -> uint32_t {
    t = t < 0 ? 0 : (t > Tlast ? Tlast : t);
    return lanebase + (uint32_t)(t - tM0);
  };
  auto fetch5 = [&](uint32_t (&d)[5], int c, int sub) {
    if (!EDGE) {
      tmem_ld_32x4(taddr + sub * L + c, d[0], d[1], d[2], d[3]);
      tmem_ld_32x1(taddr + sub * L + c + 4, d[4]);
    } else {
#pragma unroll
      for (int i = 0; i < 5; ++i) tmem_ld_32x1(colclamp(tsub + sub * L + c + i), d[i]);
    }
  };
  auto fetch6 = [&](uint32_t (&d)[6], int c, int sub) {
    if (!EDGE) {
      tmem_ld_32x4(taddr + sub * L + c, d[0], d[1], d[2], d[3]);
      tmem_ld_32x2(taddr + sub * L + c + 4, d[4], d[5]);
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) tmem_ld_32x1(colclamp(tsub + sub * L + c + i), d[i]);
    }
  };
  f32x2 X[6], V[12];
  uint32_t preA[5], preB[5], nA[6], nB[6];
  fetch5(preA, 0, 0);
  fetch5(preB, 0, 1);
  fetch6(nA, 5, 0);
  fetch6(nB, 5, 1);
  tmem_ld_wait5(preA);
  tmem_ld_wait5(preB);
#pragma unroll
  for (int i = 0; i < 5; ++i) X[i] = au_add2(pk2(__uint_as_float(preA[i]), __uint_as_float(preB[i])), bv2);
  X[5] = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 12; ++i) V[i] = pk2(0.f, 0.f);
  const int kk = ch & 63;
  const uint32_t obase = tile_u32 + (uint32_t)((ch >> 6) * p.chb + (kk & 7) * 2);
  const uint32_t och = (uint32_t)(kk >> 3);
  const int nbody = L / 6 + 1;
  for (int j = 0; j < nbody; ++j) {
    uint32_t cA[6], cB[6];
    tmem_ld_wait6(nA);
    tmem_ld_wait6(nB);
#pragma unroll
    for (int i = 0; i < 6; ++i) { cA[i] = nA[i]; cB[i] = nB[i]; }
    if (j + 1 < nbody) {
      fetch6(nA, 11 + 6 * j, 0);
      fetch6(nB, 11 + 6 * j, 1);
    }
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const f32x2 xin = au_add2(pk2(__uint_as_float(cA[s]), __uint_as_float(cB[s])), bv2);
      AU_STEP_UP(s, xin)
      if (j >= 1) {
        f32x2 y;
        AU_STEP_DOWN(s, y)
        float ya, yb;
        upk2(y, ya, yb);
        const int rowA = r0 + 6 * j + s - 6, rowB = rowA + L;
        if (EDGE) {
          const int tA = tA2 + rowA, tB = tA2 + rowB;
          if (tA < 0 || tA > Tlast) ya = 0.f;
          if (tB < 0 || tB > Tlast) yb = 0.f;
        }
        if (lane_ok) {
          st_shared_b16(obase + (uint32_t)rowA * 128u + ((och ^ (uint32_t)(rowA & 7)) << 4),
                        __bfloat16_as_ushort(__float2bfloat16_rn(ya)));
          st_shared_b16(obase + (uint32_t)rowB * 128u + ((och ^ (uint32_t)(rowB & 7)) << 4),
                        __bfloat16_as_ushort(__float2bfloat16_rn(yb)));
        }
      }
    }
  }
}

// the <= 6 a2 outputs next to the sequence ends; the warps that hold a channel group share the rows
__device__ __forceinline__ void au_a2_patch(const AUParams& p, uint32_t tile_u32, uint32_t tmem_base, int tA2, int warp,
                                            int lane) {
  const int g = warp & 3, half = warp >> 2;
  const int lane0 = (g * 32) % p.LR;
  if (lane0 >= p.Cp) return;
  const int replica = (g * 32) / p.LR;
  const int ch = lane0 + lane;
  const bool lane_ok = ch < p.Cp;
  const int T = p.T, Tlast = p.T - 1;
  const int me = replica * 2 + half, nshare = 2 * p.rep;
  const uint32_t lanebase = tmem_base + ((uint32_t)(g * 32) << 16);
  const int tM0 = tA2 - 6;
  float bv = 0.f, a = 1.f, ib = 1.f;
  if (lane_ok) {
    bv = __ldg(p.bias1 + ch);
    a = expf(__ldg(p.al2 + ch));
    ib = 1.0f / (expf(__ldg(p.be2 + ch)) + 1e-9f);
  }
  for (int q = 0; q < 6; ++q) {
    if (q % nshare != me) continue;
    const int t = q < 3 ? q : T - 6 + q;
    if (t < 0 || t > Tlast) continue;
    if (q >= 3 && t < 3) continue;
    const int row = t - tA2;
    if (row < 0 || row >= p.R2T) continue;
    uint32_t raw[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      int ti = t - 5 + i;
      ti = ti < 0 ? 0 : (ti > Tlast ? Tlast : ti);
      tmem_ld_32x1(lanebase + (uint32_t)(ti - tM0), raw[i]);
    }
    tmem_ld_wait();
    float xw[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) xw[i] = __uint_as_float(raw[i]) + bv;
    const float y = au_act_point(p.t2, a, ib, t, T, xw);
    if (lane_ok) st_shared_b16(tile_u32 + au_tile_off(row, ch, p.chb), __bfloat16_as_ushort(__float2bfloat16_rn(y)));
  }
}

// ------------------------------------------------------------------------------------------------ phase ST
// out[t0 + n] = (acc2[n] + bias + x[t0 + n]) * scale [+ accum], n < NOUT; the warps of a channel group split the columns
__device__ __forceinline__ void au_store_phase(const AUParams& p, uint32_t tmem_base, int b, int t0, int warp, int lane) {
  const int g = warp & 3, half = warp >> 2;
  const int lane0 = (g * 32) % p.LR;
  if (lane0 >= p.Cp) return;
  const int replica = (g * 32) / p.LR;
  const int ch = lane0 + lane;
  const bool lane_ok = ch < p.Cp;
  const int T = p.T;
  const int c_lo = (replica * 2 + half) * p.ncol_st;
  int c_hi = c_lo + p.ncol_st;
  if (c_hi > p.NOUT) c_hi = p.NOUT;
  if (t0 + c_hi > T) c_hi = T - t0;
  const uint32_t tbase = tmem_base + ((uint32_t)(g * 32) << 16);
  const float bv = lane_ok ? __ldg(p.bias2 + ch) : 0.f;
  const float sc = p.scale;
  const int64_t ld = p.ld;
  const int64_t base = ((int64_t)b * T + t0) * ld + ch;
  const float* xr = p.x + base;
  const float* ar = p.accum ? p.accum + base : nullptr;
  for (int c = c_lo; c < c_hi; c += 8) {
    uint32_t v[8];
    tmem_ld_32x8(tbase + (uint32_t)c, v);
    float r[8], av[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = lane_ok && c + i < c_hi;
      r[i] = ok ? BVG_LDG(xr + (int64_t)(c + i) * ld) : 0.f;
      av[i] = (ok && ar) ? BVG_LDG(ar + (int64_t)(c + i) * ld) : 0.f;
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane_ok && c + i < c_hi) {
        float y = __uint_as_float(v[i]) + bv;
        y += r[i];
        y *= sc;
        if (ar) y += av[i];
        const int64_t o = base + (int64_t)(c + i) * ld;
        if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(y);
        else reinterpret_cast<float*>(p.out)[o] = y;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(AU_THREADS, 2)
amp_unit_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                const __grid_constant__ AUParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem;
  unsigned char* tile = smem + p.nslot * p.slotb;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tile + p.nch * p.chb);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + AU_MAX_SLOTS;
  uint64_t* a1_ready = bars + 2 * AU_MAX_SLOTS;
  uint64_t* acc1_full = a1_ready + 1;
  uint64_t* a2_ready = a1_ready + 2;
  uint64_t* acc2_full = a1_ready + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a1_ready + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nslot; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(a1_ready, AU_CWARPS);
    mbar_init(acc1_full, 1);
    mbar_init(a2_ready, AU_CWARPS);
    mbar_init(acc2_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
  }
  if (warp == AU_CWARPS + 1) {
    tmem_alloc(tmem_slot, AU_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AU_CWARPS) {
    // ------------------------------------------------ weight tiles of conv1 then conv2, every tile, through the ring;
    // plus an L2 prefetch of the NEXT tile's x rows (phase A1 then reads them at L2 latency)
    const uint32_t w_bytes = (uint32_t)p.wrows * 128u;
    uint32_t ws = 0, wph = 0;
    const int span = p.NSEG1 * p.L1 + 12;
    for (int64_t tile_i = blockIdx.x; tile_i < p.n_tiles; tile_i += gridDim.x) {
      const int64_t nxt = tile_i + gridDim.x;
      if (nxt < p.n_tiles && elect_one()) {
        const int nb = (int)(nxt / p.n_ttiles);
        const int nt0 = (int)(nxt % p.n_ttiles) * p.NOUT;
        int ta = nt0 - p.h2 - 6 - p.h1 - 6;
        int tb = ta + span;
        if (ta < 0) ta = 0;
        if (tb > p.T) tb = p.T;
        if (tb > ta)
          l2_prefetch_bulk(p.x + ((int64_t)nb * p.T + ta) * p.ld, (uint32_t)((int64_t)(tb - ta) * p.ld * 4));
      }
      __syncwarp();
      for (int conv = 0; conv < 2; ++conv) {
        const CUtensorMap* tm = conv ? &tmap_w2 : &tmap_w1;
        for (int j = 0; j < p.k; ++j) {
          for (int c = 0; c < p.nch; ++c) {
            mbar_wait(&w_empty[ws], wph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&w_full[ws], w_bytes);
              tma_load_3d(ring + ws * p.slotb, tm, c * 64, 0, j, &w_full[ws]);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.nslot) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == AU_CWARPS + 1) {
    // ------------------------------------------------ MMA issuer (whole warp in the loops, one elected lane issues)
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc0 = make_smem_desc(0, 128, 0);
    const uint32_t dhi = (uint32_t)(desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)desc0 + (smem_u32(ring) >> 4), b_lo0 = (uint32_t)desc0 + (smem_u32(tile) >> 4);
    const uint32_t slot16 = (uint32_t)p.slotb >> 4, ch16 = (uint32_t)p.chb >> 4;
    uint32_t ws = 0, wph = 0, it = 0;
    for (int64_t tile_i = blockIdx.x; tile_i < p.n_tiles; tile_i += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      for (int conv = 0; conv < 2; ++conv) {
        mbar_wait(conv ? a2_ready : a1_ready, ph);
        tc_fence_after();
        const uint32_t idesc = conv ? idesc2 : idesc1;
        const uint32_t tap_step = (uint32_t)((conv ? 1 : p.dil) * 8);
        uint32_t accum = 0;
        for (int j = 0; j < p.k; ++j) {
          for (int c = 0; c < p.nch; ++c) {
            int nkk = (p.Cp - c * 64) >> 4;
            nkk = nkk > 4 ? 4 : nkk;
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + ws * slot16;
            const uint32_t b_lo = b_lo0 + (uint32_t)c * ch16 + (uint32_t)j * tap_step;
            if (elect_one()) {
              const uint64_t da = ((uint64_t)dhi << 32) | a_lo, db = ((uint64_t)dhi << 32) | b_lo;
              umma_f16_ss(tmem_base, da, db, idesc, accum);
              for (int q = 1; q < nkk; ++q) umma_f16_ss(tmem_base, da + 2 * q, db + 2 * q, idesc, 1u);
              umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            accum = 1;
            if (++ws == (uint32_t)p.nslot) { ws = 0; wph ^= 1; }
          }
        }
        if (elect_one()) umma_commit(conv ? acc2_full : acc1_full);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ compute warps 0..7
    const int tid = threadIdx.x;
    const uint32_t tile_u32 = smem_u32(tile);
    uint32_t it = 0;
    for (int64_t tile_i = blockIdx.x; tile_i < p.n_tiles; tile_i += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      const int b = (int)(tile_i / p.n_ttiles);
      const int t0 = (int)(tile_i % p.n_ttiles) * p.NOUT;
      const int tA2 = t0 - p.h2;
      const int tA1 = tA2 - 6 - p.h1;
      const bool interior = tA1 - 6 >= 0 && tA1 + p.NSEG1 * p.L1 + 5 <= p.T - 1;
      if (interior) {
        au_a1_phase<false>(p, tile_u32, b, tA1, tid);
      } else {
        au_a1_phase<true>(p, tile_u32, b, tA1, tid);
        named_bar_sync(1, AU_CTHREADS);
        au_a1_patch(p, tile_u32, b, tA1, tid);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_ready);
      mbar_wait(acc1_full, ph);
      tc_fence_after();
      if (interior) {
        au_a2_phase<false>(p, tile_u32, tmem_base, tA2, warp, lane);
      } else {
        au_a2_phase<true>(p, tile_u32, tmem_base, tA2, warp, lane);
        named_bar_sync(1, AU_CTHREADS);
        au_a2_patch(p, tile_u32, tmem_base, tA2, warp, lane);
      }
      tmem_ld_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_ready);
      mbar_wait(acc2_full, ph);
      tc_fence_after();
      au_store_phase(p, tmem_base, b, t0, warp, lane);
      tmem_ld_wait();
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AU_CWARPS + 1) tmem_dealloc(tmem_base, AU_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
static bool au_plan(const AmpUnitArgs& a, AUParams& p) {
  if (a.B <= 0 || a.T < 16 || a.T > 0x3fffffffLL) return false;
  if (a.C <= 0 || a.Cp % 16 || a.Cp < 16 || a.Cp > 96 || a.C > a.Cp || a.ld != a.Cp) return false;
  if (a.k < 1 || a.k > 11 || !(a.k & 1) || a.dil < 1) return false;
  if (!a.x || !a.out || !a.w1 || !a.w2 || !a.bias1 || !a.bias2 || !a.al1 || !a.be1 || !a.al2 || !a.be2) return false;
  if ((const void*)a.x == (const void*)a.out) return false;   // tiles read each other's halo rows
  uintptr_t al = reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.w1) |
                 reinterpret_cast<uintptr_t>(a.w2);
  if (a.accum) al |= reinterpret_cast<uintptr_t>(a.accum);
  if (al & 15) return false;
  p.x = a.x; p.out = a.out; p.accum = a.accum; p.scale = a.scale; p.out_bf16 = a.out_bf16;
  p.bias1 = a.bias1; p.bias2 = a.bias2; p.al1 = a.al1; p.be1 = a.be1; p.al2 = a.al2; p.be2 = a.be2;
  p.B = a.B; p.T = (int)a.T; p.C = a.C; p.Cp = a.Cp; p.ld = a.ld;
  p.k = a.k; p.dil = a.dil;
  p.h1 = (a.k - 1) / 2 * a.dil; p.h2 = (a.k - 1) / 2;
  p.nch = (int)ceil_div(a.Cp, 64);
  p.rep = a.Cp <= 32 ? 4 : (a.Cp <= 64 ? 2 : 1);
  p.LR = 128 / p.rep;
  if (p.rep > 1 && weight_replica_rows(a.Cp, 128) != p.LR) return false;
  p.wrows = p.rep == 1 ? round_up(a.Cp, 8) : 128;
  p.slotb = round_up(p.wrows * 128, 1024);
  p.nslot = p.nch == 2 ? 3 : 4;
  int rbmax = (AU_MAX_SMEM - 1024 - 256 - p.nslot * p.slotb) / (p.nch * 128) / 8 * 8;
  if (rbmax > 320) rbmax = 320;
  int n1max = (rbmax - 2 * p.h1) / 16 * 16;
  if (n1max > 256) n1max = 256;
  if (n1max < 64) return false;
  const int nss = 4 * p.rep;
  p.L2 = (n1max - 11) / nss / 6 * 6;
  if (p.L2 < 6) return false;
  p.R2T = nss * p.L2;
  p.N1 = round_up(p.R2T + 11, 16);
  p.N2 = (p.R2T - 2 * p.h2) / 16 * 16;
  if (p.N2 < 16) return false;
  p.NOUT = p.N2;
  p.R1 = p.N1 + 2 * p.h1;
  p.zpad = (a.Cp - a.C == 8 && a.C % 8 == 0) ? 1 : 0;
  p.P = p.zpad ? a.C / 2 : a.Cp / 2;
  int nseg = AU_CTHREADS / p.P;
  p.L1 = round_up((int)ceil_div(p.R1, nseg), 6);
  p.NSEG1 = (int)ceil_div(p.R1, p.L1);
  p.RB = round_up(p.R1 > p.R2T ? p.R1 : p.R2T, 8);
  if (p.RB > rbmax) return false;
  p.chb = round_up(p.RB * 128, 1024);
  p.ncol_st = round_up((int)ceil_div(p.NOUT, 2 * p.rep), 8);
  p.n_ttiles = (int)ceil_div(a.T, p.NOUT);
  p.n_tiles = (int64_t)a.B * p.n_ttiles;
  if (1024 + p.nslot * p.slotb + p.nch * p.chb + 256 > AU_MAX_SMEM) return false;
  return true;
}

bool amp_unit_supported(const AmpUnitArgs& a) {
  AUParams p;
  return au_plan(a, p);
}

int amp_unit_launch(const AmpUnitArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.T <= 0) return BVG_OK;
  AUParams p;
  if (!au_plan(a, p)) BVG_FAIL(BVG_EINVAL, "amp_unit: unsupported unit shape (C=%d Cp=%d k=%d dil=%d T=%lld)", a.C, a.Cp, a.k,
                               a.dil, (long long)a.T);
  make_taps_packed(&p.tp1, a.taps1);
  make_taps_packed(&p.tp2, a.taps2);
  p.t1 = a.taps1;
  p.t2 = a.taps2;
  CUtensorMap m1, m2;
  int rc = make_map_any(&m1, a.w1, 2, (uint64_t)a.Cp, 128, (uint64_t)a.k, (uint64_t)a.Cp, 64, (uint32_t)p.wrows, 1, 128);
  if (rc) return rc;
  rc = make_map_any(&m2, a.w2, 2, (uint64_t)a.Cp, 128, (uint64_t)a.k, (uint64_t)a.Cp, 64, (uint32_t)p.wrows, 1, 128);
  if (rc) return rc;
  const int smem = 1024 + p.nslot * p.slotb + p.nch * p.chb + 256;
  // kernel attributes are per device: set once for each
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0;
  BVG_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_done.load(std::memory_order_acquire) >> dev) & 1ull)) {
    BVG_CUDA(cudaFuncSetAttribute(amp_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AU_MAX_SMEM));
    BVG_CUDA(cudaFuncSetAttribute(amp_unit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (dev < 64) attr_done.fetch_or(1ull << dev, std::memory_order_release);
  }
  const int sms = umma_sm_count();
  const unsigned grid = (unsigned)(p.n_tiles < 2 * sms ? p.n_tiles : 2 * sms);
  amp_unit_kernel<<<grid, AU_THREADS, smem, st>>>(m1, m2, p);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
