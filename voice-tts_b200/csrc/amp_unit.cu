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
#include <cstring>
#include <mutex>

#include "act_packed.cuh"
#include "umma_common.cuh"

namespace bvg {

constexpr int AU_CWARPS = 8;                          // compute warps (0..7); warp 8: weight loads, warp 9: MMA issue
constexpr int AU_CTHREADS = 32 * AU_CWARPS;
constexpr int AU_THREADS = AU_CTHREADS + 64;
constexpr int AU_MAX_SMEM = 113 * 1024;               // two CTAs per SM: 2 x (113 KB + 1 KB system) = 228 KB
constexpr int AU_MAX_SLOTS = 4;
constexpr int AU_TMEM_COLS = 256;
constexpr int AU_TAIL_BYTES = 112 + 5 * 96 * 4;       // 12 barriers + TMEM slot (100 B), then the per-channel constant table

struct AUParams {
  const float* x;            // [B, T, ld] fp32
  void* out;                 // [B, T, ld] fp32 (bf16 if out_bf16)
  const float* accum;        // optional fp32 [B, T, ld]
  float scale;
  int out_bf16;
  const float *bias1, *bias2;                 // [128] fp32, zero in pad rows
  const float *al1, *be1, *al2, *be2;         // [Cp] log-scale snake parameters (pad entries 0)
  int B, T, C, Cp, ld;
  int k, dil, h1, h2;
  int nch;                   // 64-channel K chunks
  int rep, LR, wrows;
  int P, zpad, NSEG1, L1;    // phase A1: channel pairs, pad-chunk flag, segments, rows per segment
  int L2, R2T;               // phase A2: rows per sub-segment, rows per tile
  int N1, N2, NOUT, R1, RB;
  int chb;                   // bytes per 64-channel chunk of the operand tile
  int nslot, slotb;          // weight ring
  int NPst;                  // phase ST: columns per pass
  int n_ttiles;
  int64_t n_tiles;
};

// The filter taps live in constant memory (one immutable filter per device, filled on first sight by the host):
// the phase routines below are separate (non-inlined) functions, each with the uniform-register file to itself, and read
// the packed taps as uniform operands of FFMA2.  Inlined into the persistent kernel the same loops ran 80-98 instructions
// per step against 50 here: the tile loop's uniform state pushed the taps out of the uniform registers (ncu + SASS).
struct AUTaps {
  TapsPacked tp;
  Taps t;
};
__constant__ AUTaps c_au_taps;   // ONE filter per device and process (the first one seen); units with other taps run layer by layer

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
// predicated stores: skipped without a branch around the surrounding arithmetic
__device__ __forceinline__ void st_shared_b32_if(bool pred, uint32_t addr, uint32_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.b32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void st_shared_b16_if(bool pred, uint32_t addr, uint16_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.b16 [%0], %1;\n\t}" ::"r"(addr), "h"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void st_shared_zero16_if(bool pred, uint32_t addr) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.v4.b32 [%0], {%1, %1, %1, %1};\n\t}" ::"r"(addr), "r"(0u), "r"((uint32_t)pred) : "memory");
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
// thread (pair, seg): channel pair c0, c0+1; output rows [row0, row0 + 6*(nbody-1)) of the tile (time tA1 + row).
struct A1Args {
  const float* xb;       // x of this utterance, this thread's channel pair, row 0
  int ld, T;             // row pitch (elements), rows per utterance
  int tA1;               // time of tile row 0
  int row0;              // this segment's first row in the tile
  int nbody;             // L1 / 6 + 1 bodies of 6 steps (the first one produces no output)
  int nrows;             // rows >= nrows are not stored
  uint32_t obase;        // shared-memory address of this pair's bytes in tile row 0, before the 16-byte-chunk swizzle
  uint32_t och;          // the pair's 16-byte chunk inside its 128-byte row
  int zdelta;            // this thread also clears the 8 pad channels: offset from obase to their row base
  uint32_t zch;
  int zflag;
  float a0, a1, ib0, ib1;
};
template <bool EDGE, bool ZPAD>
__device__ __noinline__ void au_a1_phase(const A1Args a) {
  const TapsPacked& tp = c_au_taps.tp;
  SnakeC sn;
  sn.init(a.a0, a.a1, a.ib0, a.ib1);
  const int Tlast = a.T - 1;
  const int tf = a.tA1 + a.row0;                        // time of this segment's first output
  const int64_t ld = a.ld;
  const float* lp = a.xb + (int64_t)(tf - 6) * ld;      // interior tiles: the next row to load
  int tl = tf - 6;                                      // edge tiles: its time
  auto ldx = [&]() -> f32x2 {                           // one row of this thread's channel pair (8 bytes)
    f32x2 v;
    if (EDGE) {
      const int t = tl < 0 ? 0 : (tl > Tlast ? Tlast : tl);
      v = BVG_LDG(reinterpret_cast<const unsigned long long*>(a.xb + (int64_t)t * ld));
      ++tl;
    } else {
      v = BVG_LDG(reinterpret_cast<const unsigned long long*>(lp));
      lp += ld;
    }
    return v;
  };
  f32x2 X[6], V[12], R[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) X[i] = ldx();
#pragma unroll
  for (int i = 0; i < 6; ++i) R[i] = ldx();
  X[5] = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 12; ++i) V[i] = pk2(0.f, 0.f);
  uint32_t orow = a.obase + (uint32_t)a.row0 * 128u;    // advanced by one row per output
  int row = a.row0;
  const bool zf = ZPAD && a.zflag != 0;
  // one body = 6 steps; OUT: the steps produce outputs (all bodies but the first), LOAD: the next body's rows are fetched
#define AU_A1_BODY(OUT, LOAD)                                                                       \
  _Pragma("unroll") for (int s = 0; s < 6; ++s) {                                                   \
    const f32x2 xin = R[s];                                                                         \
    if (LOAD) R[s] = ldx();                                                                         \
    AU_STEP_UP(s, xin)                                                                              \
    if (OUT) {                                                                                      \
      f32x2 y;                                                                                      \
      AU_STEP_DOWN(s, y)                                                                            \
      float ya, yb;                                                                                 \
      upk2(y, ya, yb);                                                                              \
      if (EDGE) {                                                                                   \
        const int t = a.tA1 + row;                                                                  \
        if (t < 0 || t > Tlast) { ya = 0.f; yb = 0.f; }                                             \
      }                                                                                             \
      const uint32_t rsw = (uint32_t)(row & 7);                                                     \
      const bool ok = row < a.nrows;                                                                \
      st_shared_b32_if(ok, orow + ((a.och ^ rsw) << 4), bf16x2_bits(ya, yb));                       \
      if (ZPAD) st_shared_zero16_if(ok && zf, orow + a.zdelta + ((a.zch ^ rsw) << 4));              \
      orow += 128u;                                                                                 \
      ++row;                                                                                        \
    }                                                                                               \
  }
  AU_A1_BODY(false, true)
  for (int j = 2; j < a.nbody; ++j) AU_A1_BODY(true, true)
  AU_A1_BODY(true, false)
#undef AU_A1_BODY
}

// the <= 6 a1 outputs next to the sequence ends, one (channel, row) per thread
__device__ __noinline__ void au_a1_patch(const float* xu /*x of this utterance*/, int ld, int T, int tA1, int rows, int nc,
                                         uint32_t tile_u32, int chb, const float* ctab, int Cp, int tid) {
  const int Tlast = T - 1;
  const Taps& tp = c_au_taps.t;
  for (int item = tid; item < nc * 6; item += AU_CTHREADS) {
    const int c = item % nc, q = item / nc;
    const int t = q < 3 ? q : T - 6 + q;
    if (t < 0 || t > Tlast) continue;
    if (q >= 3 && t < 3) continue;                      // already covered by the head rows (T < 6)
    const int row = t - tA1;
    if (row < 0 || row >= rows) continue;
    float xw[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      int ti = t - 5 + i;
      ti = ti < 0 ? 0 : (ti > Tlast ? Tlast : ti);
      xw[i] = BVG_LDG(xu + (int64_t)ti * ld + c);
    }
    const float y = au_act_point(tp, ctab[c], ctab[Cp + c], t, T, xw);
    st_shared_b16(tile_u32 + au_tile_off(row, c, chb), __bfloat16_as_ushort(__float2bfloat16_rn(y)));
  }
}

// ------------------------------------------------------------------------------------------------ phase A2
// warp (lane group g, half): lanes = 32 channels; rows [r0, r0 + 2*L) of the tile as two lockstep sub-segments.
// Column n of the accumulator is time tA2 - 6 + n; a2 row r is time tA2 + r and needs columns r + 1 .. r + 11.
struct A2Args {
  uint32_t lanebase;     // TMEM address of column 0 in this warp's lane group
  int r0, L;             // first row of sub-segment A, rows per sub-segment
  int T, tA2;
  uint32_t obase;        // shared-memory address of this lane's channel in tile row 0, before the chunk swizzle
  uint32_t och;
  float a, ib, bv;
  int lane_ok;
};
template <bool EDGE>
__device__ __noinline__ void au_a2_phase(const A2Args a) {
  const TapsPacked& tp = c_au_taps.tp;
  const int L = a.L, Tlast = a.T - 1;
  const uint32_t taddr = a.lanebase + (uint32_t)a.r0;
  const int tM0 = a.tA2 - 6;                            // time of accumulator column 0
  const int tsub = a.tA2 + a.r0 - 6;                    // time of column 0 of sub-segment A (B: + L)
  SnakeC sn;
  sn.init(a.a, a.a, a.ib, a.ib);
  const f32x2 bv2 = pk2(a.bv, a.bv);
  // column fetches: c = column relative to the sub-segment's first column
  auto colclamp = [&](int t) -> uint32_t {
    t = t < 0 ? 0 : (t > Tlast ? Tlast : t);
    return a.lanebase + (uint32_t)(t - tM0);
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
  const bool lane_ok = a.lane_ok != 0;
  const int nbody = L / 6 + 1;
  uint32_t orow = a.obase + (uint32_t)a.r0 * 128u;           // sub-segment A's row of the current output (B: + L rows)
  int row = a.r0;
  const uint32_t bdelta = (uint32_t)L * 128u;
  int cnext = 11;                                            // first column of the next body's fetch
  // one body = 6 columns of both sub-segments; OUT: outputs are produced (all bodies but the first), LOAD: the next
  // body's columns are fetched while this one is computed
#define AU_A2_BODY(OUT, LOAD)                                                                        \
  {                                                                                                  \
    uint32_t cA[6], cB[6];                                                                           \
    tmem_ld_wait6(nA);                                                                               \
    tmem_ld_wait6(nB);                                                                               \
    _Pragma("unroll") for (int i = 0; i < 6; ++i) { cA[i] = nA[i]; cB[i] = nB[i]; }                  \
    if (LOAD) {                                                                                      \
      fetch6(nA, cnext, 0);                                                                          \
      fetch6(nB, cnext, 1);                                                                          \
      cnext += 6;                                                                                    \
    }                                                                                                \
    _Pragma("unroll") for (int s = 0; s < 6; ++s) {                                                  \
      const f32x2 xin = au_add2(pk2(__uint_as_float(cA[s]), __uint_as_float(cB[s])), bv2);           \
      AU_STEP_UP(s, xin)                                                                             \
      if (OUT) {                                                                                     \
        f32x2 y;                                                                                     \
        AU_STEP_DOWN(s, y)                                                                           \
        float ya, yb;                                                                                \
        upk2(y, ya, yb);                                                                             \
        const int rowB = row + L;                                                                    \
        if (EDGE) {                                                                                  \
          const int tA = a.tA2 + row, tB = a.tA2 + rowB;                                             \
          if (tA < 0 || tA > Tlast) ya = 0.f;                                                        \
          if (tB < 0 || tB > Tlast) yb = 0.f;                                                        \
        }                                                                                            \
        st_shared_b16_if(lane_ok, orow + ((a.och ^ (uint32_t)(row & 7)) << 4),                       \
                         __bfloat16_as_ushort(__float2bfloat16_rn(ya)));                             \
        st_shared_b16_if(lane_ok, orow + bdelta + ((a.och ^ (uint32_t)(rowB & 7)) << 4),             \
                         __bfloat16_as_ushort(__float2bfloat16_rn(yb)));                             \
        orow += 128u;                                                                                \
        ++row;                                                                                       \
      }                                                                                              \
    }                                                                                                \
  }
  AU_A2_BODY(false, true)
  for (int j = 2; j < nbody; ++j) AU_A2_BODY(true, true)
  AU_A2_BODY(true, false)
#undef AU_A2_BODY
}

// the <= 6 a2 outputs next to the sequence ends; the `nshare` warps that hold a channel group share the rows
__device__ __noinline__ void au_a2_patch(uint32_t lanebase, int T, int tA2, int R2T, int me, int nshare, uint32_t tile_u32,
                                         int chb, int ch, int lane_ok, float a, float ib, float bv) {
  const int Tlast = T - 1;
  const int tM0 = tA2 - 6;
  const Taps& tp = c_au_taps.t;
  for (int q = 0; q < 6; ++q) {
    if (q % nshare != me) continue;
    const int t = q < 3 ? q : T - 6 + q;
    if (t < 0 || t > Tlast) continue;
    if (q >= 3 && t < 3) continue;
    const int row = t - tA2;
    if (row < 0 || row >= R2T) continue;
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
    const float y = au_act_point(tp, a, ib, t, T, xw);
    if (lane_ok) st_shared_b16(tile_u32 + au_tile_off(row, ch, chb), __bfloat16_as_ushort(__float2bfloat16_rn(y)));
  }
}

// ------------------------------------------------------------------------------------------------ phase ST
// out[t0 + n] = (acc2[n] + bias + x[t0 + n]) * scale [+ accum], n < NOUT.  With row pitch == Cp the tile's rows of x / out /
// accum are ONE contiguous block of NV * Cp floats, so the phase is a contiguous float4 stream over all 256 threads:
//   the residual block is loaded into registers BEFORE the wait for conv2 (it does not depend on it);
//   the warps that hold a channel group copy their accumulator columns TMEM -> shared memory [column][Cp] fp32 (the
//   operand tile is free once conv2 has completed) - the transposition that makes the stream contiguous;
//   every thread then adds bias / residual and stores 16 bytes per instruction.
// Tiles whose fp32 block exceeds the operand tile's bytes run in passes of NP columns.
constexpr int AU_ST_ITEMS = 11;        // float4 items per thread and pass (256 threads: <= 45 KB per pass)
struct STArgs {
  const float* xblk;     // x rows of this tile (block start)
  void* oblk;            // out rows (fp32 or bf16 elements)
  const float* ablk;     // accum rows or nullptr
  int n4;                // float4 items in the tile = valid rows * Cp / 4
  int np4;               // float4 items per pass = NP * Cp / 4
  int NP, NV;            // columns per pass, valid columns of the tile
  int Cp;
  uint32_t stg;          // shared-memory staging (the operand tile)
  uint32_t bias_s;       // bias2[Cp] fp32 in shared memory
  float sc;
  int obf;
  uint32_t lanebase;     // TMEM lane group of this warp
  int ch, lane_ok, warp_ok;
  int me, nshare;        // this warp's index among the warps that hold its channel group
  uint64_t* bar;         // conv2-complete barrier and its parity for this tile
  uint32_t ph;
  int tid;
};
__device__ __noinline__ void au_store_phase(const STArgs a) {
  const int q4 = a.Cp >> 2;                              // float4 per row
  const float4* xb = reinterpret_cast<const float4*>(a.xblk);
  const float4* ab = reinterpret_cast<const float4*>(a.ablk);
  const int qstep = AU_CTHREADS % q4;
  for (int i0 = 0, pass = 0; i0 < a.n4; i0 += a.np4, ++pass) {
    int iend = i0 + a.np4;
    if (iend > a.n4) iend = a.n4;
    // residual rows of this pass (x * scale + accum when the epilogue has them: (s + b + x) * sc + acc = (s + b) * sc + that)
    float4 xa[AU_ST_ITEMS];
#pragma unroll
    for (int m = 0; m < AU_ST_ITEMS; ++m) {
      const int i = i0 + a.tid + m * AU_CTHREADS;
      xa[m] = i < iend ? BVG_LDG(xb + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (pass == 0) {
      if (a.tid < 32) mbar_wait_sleep(a.bar, a.ph, 40);    // one warp watches the mbarrier, the others park on a hardware barrier
      named_bar_sync(2, AU_CTHREADS);
      tc_fence_after();
    }
    // accumulator columns [c_lo, c_hi) of this pass -> staging, chunks of 8 columns interleaved over the sharing warps
    if (a.warp_ok) {
      const int c_lo = pass * a.NP;
      int c_hi = c_lo + a.NP;
      if (c_hi > a.NV) c_hi = a.NV;
      for (int c = c_lo + a.me * 8; c < c_hi; c += 8 * a.nshare) {
        uint32_t v[8];
        tmem_ld_32x8(a.lanebase + (uint32_t)c, v);
        tmem_ld_wait();
        const uint32_t o = a.stg + (uint32_t)(((c - c_lo) * a.Cp + a.ch) * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (a.lane_ok && c + i < c_hi) st_shared_f32(o + (uint32_t)(i * a.Cp * 4), __uint_as_float(v[i]));
      }
    }
    if (a.ablk) {
#pragma unroll
      for (int m = 0; m < AU_ST_ITEMS; ++m) {
        const int i = i0 + a.tid + m * AU_CTHREADS;
        if (i < iend) {
          const float4 c = BVG_LDG(ab + i);
          xa[m].x = fmaf(xa[m].x, a.sc, c.x); xa[m].y = fmaf(xa[m].y, a.sc, c.y);
          xa[m].z = fmaf(xa[m].z, a.sc, c.z); xa[m].w = fmaf(xa[m].w, a.sc, c.w);
        }
      }
    } else if (a.sc != 1.f) {
#pragma unroll
      for (int m = 0; m < AU_ST_ITEMS; ++m) { xa[m].x *= a.sc; xa[m].y *= a.sc; xa[m].z *= a.sc; xa[m].w *= a.sc; }
    }
    named_bar_sync(1, AU_CTHREADS);
    int q = a.tid % q4;
#pragma unroll
    for (int m = 0; m < AU_ST_ITEMS; ++m) {
      const int i = i0 + a.tid + m * AU_CTHREADS;
      if (i < iend) {
        float4 sv, bv;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w)
                     : "r"(a.stg + (uint32_t)((i - i0) * 16)));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bv.x), "=f"(bv.y), "=f"(bv.z), "=f"(bv.w)
                     : "r"(a.bias_s + (uint32_t)(q * 16)));
        float4 y;
        if (a.sc == 1.f && !a.ablk) {                      // the plain unit: (acc + bias) + x, the order of the conv epilogues
          y.x = (sv.x + bv.x) + xa[m].x; y.y = (sv.y + bv.y) + xa[m].y;
          y.z = (sv.z + bv.z) + xa[m].z; y.w = (sv.w + bv.w) + xa[m].w;
        } else {
          y.x = fmaf(sv.x + bv.x, a.sc, xa[m].x); y.y = fmaf(sv.y + bv.y, a.sc, xa[m].y);
          y.z = fmaf(sv.z + bv.z, a.sc, xa[m].z); y.w = fmaf(sv.w + bv.w, a.sc, xa[m].w);
        }
        if (a.obf) {
          uint2 o;
          o.x = bf16x2_bits(y.x, y.y);
          o.y = bf16x2_bits(y.z, y.w);
          reinterpret_cast<uint2*>(a.oblk)[i] = o;
        } else {
          reinterpret_cast<float4*>(a.oblk)[i] = y;
        }
      }
      q += qstep;
      if (q >= q4) q -= q4;
    }
    named_bar_sync(1, AU_CTHREADS);                       // staging is rewritten by the next pass / the next tile's phase A1
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
  // per-channel constants: [0] exp(alpha1), [1] 1/(exp(beta1)+1e-9), [2] / [3] the same of the second activation, [4] bias2
  float* ctab = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 112);

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
  for (int c = threadIdx.x; c < p.Cp; c += AU_THREADS) {
    ctab[c] = expf(__ldg(p.al1 + c));
    ctab[p.Cp + c] = 1.0f / (expf(__ldg(p.be1 + c)) + 1e-9f);
    ctab[2 * p.Cp + c] = expf(__ldg(p.al2 + c));
    ctab[3 * p.Cp + c] = 1.0f / (expf(__ldg(p.be2 + c)) + 1e-9f);
    ctab[4 * p.Cp + c] = __ldg(p.bias2 + c);
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
            mbar_wait_sleep(&w_empty[ws], wph ^ 1, 100);
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
        mbar_wait_sleep(conv ? a2_ready : a1_ready, ph, 60);
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
    // tile-invariant: the phase-A1 (pair, segment) of this thread, the lane's channel and role in the TMEM phases
    const int seg1 = tid / p.P, pair1 = tid - seg1 * p.P;
    const int g = warp & 3, half = warp >> 2;
    const int lane0 = (g * 32) % p.LR, replica = (g * 32) / p.LR;
    const int chT = lane0 + lane;
    const bool warp_ok = lane0 < p.Cp;
    const bool lane_ok = warp_ok && chT < p.Cp;
    const int segT = replica * 2 + half;                 // which share of a channel group's rows / columns this warp takes
    const float bv1 = lane_ok ? __ldg(p.bias1 + chT) : 0.f;
    const uint32_t lanebase = tmem_base + ((uint32_t)(g * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile_i = blockIdx.x; tile_i < p.n_tiles; tile_i += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      const int b = (int)(tile_i / p.n_ttiles);
      const int t0 = (int)(tile_i % p.n_ttiles) * p.NOUT;
      const int tA2 = t0 - p.h2;
      const int tA1 = tA2 - 6 - p.h1;
      const bool interior = tA1 - 6 >= 0 && tA1 + p.NSEG1 * p.L1 + 5 <= p.T - 1;
      const float* xu = p.x + (int64_t)b * p.T * p.ld;
      if (seg1 < p.NSEG1) {
        const int c0 = 2 * pair1, kk = c0 & 63;
        A1Args a;
        a.xb = xu + c0; a.ld = p.ld; a.T = p.T; a.tA1 = tA1; a.row0 = seg1 * p.L1;
        a.nbody = p.L1 / 6 + 1; a.nrows = p.RB;
        a.obase = tile_u32 + (uint32_t)((c0 >> 6) * p.chb + (kk & 7) * 2);
        a.och = (uint32_t)(kk >> 3);
        a.zflag = p.zpad && pair1 == p.P - 1;
        a.zdelta = (p.C >> 6) * p.chb - ((c0 >> 6) * p.chb + (kk & 7) * 2);
        a.zch = (uint32_t)((p.C & 63) >> 3);
        a.a0 = ctab[c0]; a.a1 = ctab[c0 + 1]; a.ib0 = ctab[p.Cp + c0]; a.ib1 = ctab[p.Cp + c0 + 1];
        if (p.zpad) {
          if (interior) au_a1_phase<false, true>(a);
          else au_a1_phase<true, true>(a);
        } else {
          if (interior) au_a1_phase<false, false>(a);
          else au_a1_phase<true, false>(a);
        }
      }
      if (!interior) {
        named_bar_sync(1, AU_CTHREADS);
        const int rows = p.NSEG1 * p.L1 < p.RB ? p.NSEG1 * p.L1 : p.RB;
        au_a1_patch(xu, p.ld, p.T, tA1, rows, 2 * p.P, tile_u32, p.chb, ctab, p.Cp, tid);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_ready);
      // conv1 complete: ONE warp watches the mbarrier, the others park on a hardware barrier (no issue slots spent)
      if (warp == 0) mbar_wait_sleep(acc1_full, ph, 40);
      named_bar_sync(2, AU_CTHREADS);
      tc_fence_after();
      const float a2a = lane_ok ? ctab[2 * p.Cp + chT] : 1.f, a2ib = lane_ok ? ctab[3 * p.Cp + chT] : 1.f;
      if (warp_ok) {
        const int kk = chT & 63;
        A2Args a;
        a.lanebase = lanebase; a.r0 = segT * 2 * p.L2; a.L = p.L2; a.T = p.T; a.tA2 = tA2;
        a.obase = tile_u32 + (uint32_t)((chT >> 6) * p.chb + (kk & 7) * 2);
        a.och = (uint32_t)(kk >> 3);
        a.a = a2a; a.ib = a2ib; a.bv = bv1; a.lane_ok = lane_ok;
        if (interior) au_a2_phase<false>(a);
        else au_a2_phase<true>(a);
      }
      if (!interior) {
        named_bar_sync(1, AU_CTHREADS);
        if (warp_ok)
          au_a2_patch(lanebase, p.T, tA2, p.R2T, segT, 2 * p.rep, tile_u32, p.chb, chT, lane_ok, a2a, a2ib, bv1);
      }
      tmem_ld_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_ready);
      // phase ST (waits for conv2 inside, after its residual loads are in flight)
      {
        int nv = p.T - t0;
        if (nv > p.NOUT) nv = p.NOUT;
        const int64_t o0 = ((int64_t)b * p.T + t0) * p.ld;
        STArgs a;
        a.xblk = p.x + o0;
        a.oblk = p.out_bf16 ? (void*)(reinterpret_cast<__nv_bfloat16*>(p.out) + o0) : (void*)(reinterpret_cast<float*>(p.out) + o0);
        a.ablk = p.accum ? p.accum + o0 : nullptr;
        a.n4 = nv * (p.Cp >> 2); a.np4 = p.NPst * (p.Cp >> 2); a.NP = p.NPst; a.NV = nv; a.Cp = p.Cp;
        a.stg = tile_u32; a.bias_s = smem_u32(ctab + 4 * p.Cp); a.sc = p.scale; a.obf = p.out_bf16;
        a.lanebase = lanebase; a.ch = chT; a.lane_ok = lane_ok; a.warp_ok = warp_ok; a.me = segT; a.nshare = 2 * p.rep;
        a.bar = acc2_full; a.ph = ph; a.tid = tid;
        au_store_phase(a);
      }
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
  if (a.B <= 0 || a.T < 16 || a.T > 0x3fffffffLL || a.T * a.Cp > 0x7fffffffLL) return false;
  if (a.Cp % 4) return false;
  if (a.C <= 0 || a.Cp % 16 || a.Cp < 16 || a.Cp > 96 || a.C > a.Cp || a.ld != a.Cp) return false;
  if (a.k < 1 || a.k > 11 || !(a.k & 1) || a.dil < 1) return false;
  if (!a.x || !a.out || !a.w1 || !a.w2 || !a.bias1 || !a.bias2 || !a.al1 || !a.be1 || !a.al2 || !a.be2) return false;
  if ((const void*)a.x == (const void*)a.out) return false;   // tiles read each other's halo rows
  if (memcmp(&a.taps1, &a.taps2, sizeof(Taps)) != 0) return false;   // one filter for both activations
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
  int rbmax = (AU_MAX_SMEM - 1024 - AU_TAIL_BYTES - p.nslot * p.slotb) / (p.nch * 128) / 8 * 8;
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
  {
    // phase ST stages the tile's fp32 block in the operand tile: passes of NPst columns, <= AU_ST_ITEMS float4 per thread
    int cap = p.nch * p.chb;
    if (cap > AU_ST_ITEMS * AU_CTHREADS * 16) cap = AU_ST_ITEMS * AU_CTHREADS * 16;
    const int npass = (int)ceil_div((int64_t)p.NOUT * a.Cp * 4, cap);
    p.NPst = round_up((int)ceil_div(p.NOUT, npass), 8);
    if ((int64_t)p.NPst * a.Cp * 4 > cap) return false;
  }
  p.n_ttiles = (int)ceil_div(a.T, p.NOUT);
  p.n_tiles = (int64_t)a.B * p.n_ttiles;
  if (1024 + p.nslot * p.slotb + p.nch * p.chb + AU_TAIL_BYTES > AU_MAX_SMEM) return false;
  return true;
}

// true when `t` is (or now becomes) the filter held in c_au_taps on the current device.  The constant is written once per
// device, before its first use, and never changed afterwards, so launches on any stream may read it.
static bool au_taps_resident(const Taps& t) {
  static std::mutex mu;
  static Taps known[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  std::lock_guard<std::mutex> lk(mu);
  if (have[dev]) return memcmp(&known[dev], &t, sizeof(Taps)) == 0;
  AUTaps h;
  make_taps_packed(&h.tp, t);
  h.t = t;
  if (cudaMemcpyToSymbol(c_au_taps, &h, sizeof(h), 0, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  known[dev] = t;
  have[dev] = true;
  return true;
}

bool amp_unit_supported(const AmpUnitArgs& a) {
  AUParams p;
  return au_plan(a, p) && au_taps_resident(a.taps1);
}

int amp_unit_launch(const AmpUnitArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.T <= 0) return BVG_OK;
  AUParams p;
  if (!au_plan(a, p)) BVG_FAIL(BVG_EINVAL, "amp_unit: unsupported unit shape (C=%d Cp=%d k=%d dil=%d T=%lld)", a.C, a.Cp, a.k,
                               a.dil, (long long)a.T);
  if (!au_taps_resident(a.taps1)) BVG_FAIL(BVG_EINVAL, "amp_unit: a different anti-aliasing filter is resident on this device");
  CUtensorMap m1, m2;
  int rc = make_map_any(&m1, a.w1, 2, (uint64_t)a.Cp, 128, (uint64_t)a.k, (uint64_t)a.Cp, 64, (uint32_t)p.wrows, 1, 128);
  if (rc) return rc;
  rc = make_map_any(&m2, a.w2, 2, (uint64_t)a.Cp, 128, (uint64_t)a.k, (uint64_t)a.Cp, 64, (uint32_t)p.wrows, 1, 128);
  if (rc) return rc;
  const int smem = 1024 + p.nslot * p.slotb + p.nch * p.chb + AU_TAIL_BYTES;
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
