// Dilated Conv1d with the FOLLOWING anti-aliased activation fused into its epilogue:
//
//   out[b, t, co] = bf16( Activation1d_{alpha,beta}( conv1d(in)[b, :, co] + bias[co] )[t] )
//
// i.e. `xt = c1(xt); xt = a2(xt)` of AMPBlock1.forward (bigvgan.py:132-141) in one kernel: the fp32
// convolution result never leaves the SM (no bf16 rounding of the intermediate, no HBM round trip, no
// separate activation launch).  The implicit GEMM (TMA producers, single-thread tcgen05.mma issue,
// 128-byte operand rows, resident / streamed weights, replicated weight rows for narrow layers) is the
// one of conv_umma2.cu; what differs is the tile geometry and the epilogue.
//
// Channel-major accumulators put one output channel on each TMEM lane and TIME on the columns, so the
// FIR runs along the registers of a thread - the same sliding-window scheme as act1d_cl.cu:
//   u[2t]   = sum_q up[2q+1] * x[clamp(t+2-q)],  u[2t+1] = sum_q up[2q] * x[clamp(t+3-q)]
//   v[m]    = u[m] + 1/(b+1e-9) * sin(a*u[m])^2
//   y[t]    = sum_k down[k] * v[clamp(2t+k-5, 0, 2T-1)]
// (alias_free_activation/torch/{resample.py:29-38,55-58, filter.py:94-101, act.py:25-30}, activations.py:107-120),
// x being the convolution output.  y[t] needs x[t-5 .. t+5], so a tile computes NT = 256 convolution
// columns (times t0-6 .. t0+249) and emits NOUT = 240 activation outputs (times t0 .. t0+239): tiles
// overlap by 16 columns (6.25 % extra MMA work; 208 / 192 for <= 32-channel layers).  Sequence ends follow
// the torch operator (replicate padding of x, then of v), applied in registers by the edge variant of the
// segment routine.
//
// Warps (12): 0 activation-tile loads | 1 TMEM alloc + MMA issue | 2-9 epilogue | 10,11 weight loads.
// Epilogue: the 2*rep warps that hold the same 32 channels (two per TMEM lane group, times the weight-row
// replicas of narrow layers) split the outputs into contiguous segments of S = 120 / 60 / 24.  A warp walks
// its segment as two lockstep sub-segments packed in f32x2 registers, in bodies of 6 columns (TMEM loads of
// the next body in flight while the current one is computed), stages 2 x 30 output rows x 32 channels (bf16)
// in a private double buffer and hands them to its own TMA tensor stores - there is no cross-warp
// synchronisation in the epilogue at all.  (Storing the row slices straight from registers - 64 contiguous
// bytes per warp-wide store - was measured 10-20 % slower on the narrow layers.)
#include <cstdlib>

#include "act_packed.cuh"
#include "umma_common.cuh"

namespace bvg {

constexpr int UA_EPI_WARPS = 8;
constexpr int UA_WARPS = 12;
constexpr int UA_THREADS = 32 * UA_WARPS;
constexpr int UA_SLOT_BYTES = 16384;                 // one weight stage (128 rows x 128 B)
constexpr int UA_A_SLOTS = 5;
constexpr int UA_X_STAGES = 2;
constexpr int UA_MAX_X_ROWS = 320;
constexpr int UA_X_STAGE_BYTES = UA_MAX_X_ROWS * 128;
constexpr int UA_LEAD = 6;                           // columns before the first output: 5 halo + 1 (bodies of 6 steps)
constexpr int UA_STAGE_BYTES = 2 * 30 * 32 * 2;      // one staging buffer: 2 sub-segments x 30 rows x 32 channels x bf16
constexpr int UA_SMEM_USED = 1024 + UA_A_SLOTS * UA_SLOT_BYTES + UA_X_STAGES * UA_X_STAGE_BYTES +
                             UA_EPI_WARPS * 2 * UA_STAGE_BYTES + 512;
constexpr int UA_SMEM_BYTES = 227 * 1024;            // the whole opt-in maximum: this CTA owns the SM (see conv_umma2.cu)
static_assert(UA_SMEM_USED <= UA_SMEM_BYTES, "shared-memory plan exceeds the opt-in maximum");

struct UAParams {
  const float* bias;
  const float* alpha_log;   // [Cout] log-scale snake parameters of the fused activation
  const float* beta_log;
  const float* res;         // RES mode: fp32 residual [B, T, out_ld] ...
  float* y;                 // ... and the fp32 sum (conv + bias + res) written for the next unit's residual
  int out_ld;
  TapsPacked tp;
  int B, T;
  int Cin_p, nchunks;
  int k, dil, center;
  int x_box_rows, x_nbox;
  int n_ttiles, n_cotiles;
  int64_t n_tiles;
  int CW;            // out channels per tile
  int rep, LR;       // weight replicas in the 128 MMA rows, lanes per replica
  int wrows;
  int a_stages;
  int w_resident;
  int S;             // outputs per warp segment (two lockstep sub-segments of S/2): 120 / 60 / 24 for rep 1 / 2 / 4
  int Rs;            // rows per sub-segment per staging round (TMA store box rows): 30 / 30 / 12
  int NOUT;          // activation outputs per tile = 2 * rep * S (240 / 240 / 192)
  int NT;            // convolution columns per tile (MMA N) >= NOUT + 11: 256 / 256 / 208
};

struct UATile {
  int cot, b, t0;
};
__device__ __forceinline__ UATile ua_tile(const UAParams& p, int64_t tile) {
  UATile t;
  t.cot = (int)(tile % p.n_cotiles);
  const int64_t r = tile / p.n_cotiles;
  t.t0 = (int)(r % p.n_ttiles) * p.NOUT;
  t.b = (int)(r / p.n_ttiles);
  return t;
}

__device__ __forceinline__ void ua_tma_store(const void* tmap, uint32_t saddr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(saddr), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// One segment of a warp: outputs y[tseg .. tseg+S) of its 32 channels from TMEM columns [taddr, taddr+S+11).
// The segment is walked as TWO sub-segments of L = S/2 outputs in lockstep, packed into the two halves of
// f32x2 registers (lo = A: outputs tseg.., hi = B: outputs tseg+L..): FFMA2 halves the issue count and the two
// independent recurrences give the two epilogue warps of a scheduler enough work to cover the FMA latency.
// Step i of a sub-segment (t = t_first - 6 + i) takes x[t+5] into the window and produces v[2t+5], v[2t+6]
// and (i >= 6) y[t].  EDGE: the segment touches a sequence end (x / v replicate rules, as the generic path of
// act1d_cl_kernel); handled per half on unpacked values.
// RES: x = accumulator + bias + res[t][channel] (the AMPBlock1 residual, bigvgan.py:139); x is also stored as the fp32
// residual stream y[t][channel] for the rows this sub-segment owns.  rrow / yrow point at column 0 of sub-segment A
// (time tseg - 6) / at its first owned row (time tseg) for this lane's channel; rows are p.out_ld elements apart.
template <bool EDGE, bool RES>
__device__ __forceinline__ void ua_segment(const UAParams& p, uint32_t taddr, int tseg, float bv, float a, float ib,
                                           uint32_t stg, int wbox, int lane, bool lane_ok, const void* omap,
                                           int ch0, int b, uint32_t& nstore, const float* __restrict__ rrow,
                                           float* __restrict__ yrow) {
  const int L = p.S >> 1, Rs = p.Rs, T = p.T, Tlast = p.T - 1;
  const int64_t ld = p.out_ld;
  // residual row of column c (time tseg - 6 + c); sequence ends: any in-range row will do, the value is overridden
  auto rload = [&](int c) -> float {
    if (!RES) return 0.f;
    int t = tseg - 6 + c;
    if (EDGE) t = t < 0 ? 0 : (t > Tlast ? Tlast : t);
    return lane_ok ? BVG_LDG(rrow + (int64_t)(t - (tseg - 6)) * ld) : 0.f;
  };
  const int bodies_per_round = Rs / 6;
  const uint32_t half_bytes = UA_STAGE_BYTES / 2;            // A rows, then (128-byte aligned for the TMA) B rows
  const float hbf = 0.5f * ib;
  const f32x2 hb = pk2(hbf, hbf), nhb = pk2(-hbf, -hbf), a2 = pk2(2.0f * a, 2.0f * a), na2hb = pk2(-2.0f * a * hbf, -2.0f * a * hbf);
  const f32x2 bv2 = pk2(bv, bv);
  f32x2 X[6], V[12];
  uint32_t preA[5], preB[5], nA[6], nB[6];
  tmem_ld_32x4(taddr, preA[0], preA[1], preA[2], preA[3]);
  tmem_ld_32x1(taddr + 4, preA[4]);
  tmem_ld_32x4(taddr + L, preB[0], preB[1], preB[2], preB[3]);
  tmem_ld_32x1(taddr + L + 4, preB[4]);
  tmem_ld_32x4(taddr + 5, nA[0], nA[1], nA[2], nA[3]);
  tmem_ld_32x2(taddr + 9, nA[4], nA[5]);
  tmem_ld_32x4(taddr + L + 5, nB[0], nB[1], nB[2], nB[3]);
  tmem_ld_32x2(taddr + L + 9, nB[4], nB[5]);
  float rA[6], rB[6];                           // residual of the current body; slot s is refilled for the next body once used
  {
    float qA[5], qB[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { qA[i] = rload(i); qB[i] = rload(L + i); }
#pragma unroll
    for (int i = 0; i < 6; ++i) { rA[i] = rload(5 + i); rB[i] = rload(L + 5 + i); }
    tmem_ld_wait5(preA);
    tmem_ld_wait5(preB);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      X[i] = add2(pk2(__uint_as_float(preA[i]), __uint_as_float(preB[i])), bv2);
      if (RES) X[i] = add2(X[i], pk2(qA[i], qB[i]));
    }
  }
  X[5] = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 12; ++i) V[i] = pk2(0.f, 0.f);
  float xlastA, xlastB, vendA = 0.f, vendB = 0.f;
  upk2(X[4], xlastA, xlastB);
  int nbody = L / 6 + 1;
  if (EDGE) {
    // rounds whose first row (of sub-segment A) is past the end of the sequence are not computed
    const int rounds = (T - tseg + Rs - 1) / Rs;
    if (rounds * bodies_per_round + 1 < nbody) nbody = rounds * bodies_per_round + 1;
  }
  uint32_t obuf = 0;
  int rpos = 0, round = 0;                      // body position inside its staging round, round index
  for (int j = 0; j < nbody; ++j) {
    uint32_t cA[6], cB[6];
    tmem_ld_wait6(nA);
    tmem_ld_wait6(nB);
#pragma unroll
    for (int i = 0; i < 6; ++i) { cA[i] = nA[i]; cB[i] = nB[i]; }
    if (j + 1 < nbody) {
      tmem_ld_32x4(taddr + 11 + 6 * j, nA[0], nA[1], nA[2], nA[3]);
      tmem_ld_32x2(taddr + 15 + 6 * j, nA[4], nA[5]);
      tmem_ld_32x4(taddr + L + 11 + 6 * j, nB[0], nB[1], nB[2], nB[3]);
      tmem_ld_32x2(taddr + L + 15 + 6 * j, nB[4], nB[5]);
    }
    if (j >= 1 && rpos == 0) {
      // the buffer about to be written was handed to the TMA two rounds ago
      if (elect_one()) bulk_wait_group_read<1>();   // elect.sync is deterministic: the same lane owns all bulk groups
      __syncwarp();
      obuf = stg + (nstore & 1u) * UA_STAGE_BYTES;
    }
    const bool left = EDGE && j == 0 && tseg == 0;
    if (left) {
      // sub-segment A starts the sequence: x[t < 0] := x[0]; x[0] is the second column of this body
      const float x0 = __uint_as_float(cA[1]) + bv + (RES ? rA[1] : 0.f);
#pragma unroll
      for (int i = 0; i < 5; ++i) { float lo, hi; upk2(X[i], lo, hi); X[i] = pk2(x0, hi); }
      cA[0] = cA[1];
      if (RES) rA[0] = rA[1];
      xlastA = x0;
    }
    const int tb = tseg - 6 + 6 * j;           // t of step 0 of this body in sub-segment A (B: + L)
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      f32x2 xin = add2(pk2(__uint_as_float(cA[s]), __uint_as_float(cB[s])), bv2);
      if (RES) {
        xin = add2(xin, pk2(rA[s], rB[s]));
        if (j + 1 < nbody) { rA[s] = rload(11 + 6 * j + s); rB[s] = rload(L + 11 + 6 * j + s); }
        // this column is time tseg + o (A) / tseg + L + o (B); rows 0 <= o < L are the ones the sub-segment owns
        const int o = 6 * j + s - 1;
        if (lane_ok && o >= 0 && o < L) {
          float ya, yb;
          upk2(xin, ya, yb);
          if (!EDGE || tseg + o <= Tlast) yrow[(int64_t)o * ld] = ya;
          if (!EDGE || tseg + L + o <= Tlast) yrow[(int64_t)(L + o) * ld] = yb;
        }
      }
      if (EDGE) {
        float xa, xb;
        upk2(xin, xa, xb);
        if (tb + s + 5 > Tlast) xa = xlastA; else xlastA = xa;
        if (tb + L + s + 5 > Tlast) xb = xlastB; else xlastB = xb;
        xin = pk2(xa, xb);
      }
      X[(s + 5) % 6] = xin;
      f32x2 uo = hb, ue = hb;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const f32x2 xv = X[(s + 5 - q) % 6];
        uo = fma2(p.tp.u[q], xv, uo);
        ue = fma2(p.tp.u[5 - q], xv, ue);
      }
      float zo0, zo1, ze0, ze1;
      upk2(fma2(a2, uo, na2hb), zo0, zo1);
      upk2(fma2(a2, ue, na2hb), ze0, ze1);
      f32x2 vo = fma2(nhb, pk2(__cosf(zo0), __cosf(zo1)), uo);
      f32x2 ve = fma2(nhb, pk2(__cosf(ze0), __cosf(ze1)), ue);
      if (EDGE) {
        float voA, voB, veA, veB;
        upk2(vo, voA, voB);
        upk2(ve, veA, veB);
        const int tA = tb + s, tB = tb + L + s;
        if (tA >= T - 3) {                     // v[m >= 2T] := v[2T-1], the odd sample of step T-3
          if (tA == T - 3) vendA = voA;
          voA = vendA; veA = vendA;
        }
        if (tB >= T - 3) {
          if (tB == T - 3) vendB = voB;
          voB = vendB; veB = vendB;
        }
        vo = pk2(voA, voB);
        ve = pk2(veA, veB);
      }
      V[(2 * s + 10) % 12] = vo;
      V[(2 * s + 11) % 12] = ve;
      if (left && s == 3) {                    // v[m < 0] := v[0], the even sample of step t = -3 (sub-segment A only)
        float v0, hi;
        upk2(V[5], v0, hi);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const int slot = (10 + i) % 12;      // 10, 11, 0, 1, 2, 3, 4
          float lo, h2;
          upk2(V[slot], lo, h2);
          V[slot] = pk2(v0, h2);
        }
      }
      if (j >= 1) {
        // 12-tap decimating FIR as two independent 6-term chains (even / odd taps)
        f32x2 ae = mul2(p.tp.d[0], V[(2 * s) % 12]);
        f32x2 ao = mul2(p.tp.d[1], V[(2 * s + 1) % 12]);
#pragma unroll
        for (int k = 2; k < 12; k += 2) {
          ae = fma2(p.tp.d[k < 6 ? k : 11 - k], V[(2 * s + k) % 12], ae);
          ao = fma2(p.tp.d[k + 1 < 6 ? k + 1 : 10 - k], V[(2 * s + k + 1) % 12], ao);
        }
        float ya, yb;
        upk2(add2(ae, ao), ya, yb);
        if (lane_ok) {
          const uint32_t o = obuf + (uint32_t)(((rpos * 6 + s) * wbox + lane) * 2);
          st_shared_b16(o, __bfloat16_as_ushort(__float2bfloat16_rn(ya)));
          st_shared_b16(o + half_bytes, __bfloat16_as_ushort(__float2bfloat16_rn(yb)));
        }
      }
    }
    if (j >= 1) {
      if (++rpos == bodies_per_round) {
        rpos = 0;
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          const int rowA = tseg + round * Rs;    // rows >= T are clipped by the tensor map
          ua_tma_store(omap, obuf, ch0, rowA, b);
          if (rowA + L < T) ua_tma_store(omap, obuf + half_bytes, ch0, rowA + L, b);
          bulk_commit_group();
        }
        ++nstore;
        ++round;
      }
    }
  }
}

template <bool RES>
__global__ void __launch_bounds__(UA_THREADS, 1)
conv_umma2a_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out_tail,
                   const __grid_constant__ UAParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_st = smem;
  unsigned char* x_st = smem + UA_A_SLOTS * UA_SLOT_BYTES;
  unsigned char* o_st = x_st + UA_X_STAGES * UA_X_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_st + UA_EPI_WARPS * 2 * UA_STAGE_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + UA_A_SLOTS;
  uint64_t* x_full = a_empty + UA_A_SLOTS;
  uint64_t* x_empty = x_full + UA_X_STAGES;
  uint64_t* t_full = x_empty + UA_X_STAGES;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < UA_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < UA_X_STAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], UA_EPI_WARPS); }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    tma_prefetch_desc(&tmap_out_tail);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // no global access above this line (common.cuh, PDL)
  const uint32_t tmem_base = *tmem_slot;
  const int CW = p.CW;

  if (warp == 0) {
    {
      // ------------------------------------------------ activation tiles (rows t0-6-center*dil ...; TMA zero fill = conv padding)
      const uint32_t xbox_bytes = (uint32_t)p.x_box_rows * 128u;
      uint32_t xs = 0, xph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const UATile t = ua_tile(p, tile);
        const int trow = t.t0 - UA_LEAD - p.center * p.dil;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&x_empty[xs], xph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&x_full[xs], (uint32_t)p.x_nbox * xbox_bytes);
            unsigned char* dstx = x_st + xs * UA_X_STAGE_BYTES;
            for (int q = 0; q < p.x_nbox; ++q)
              tma_load_3d(dstx + q * xbox_bytes, &tmap_x, c * 64, trow + q * p.x_box_rows, t.b, &x_full[xs]);
          }
          __syncwarp();
          if (++xs == UA_X_STAGES) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == 10 || warp == 11) {
    {
      // ------------------------------------------------ weight tiles: this warp owns every second ring stage
      const uint32_t mine = (uint32_t)(warp - 10);
      const uint32_t a_bytes = (uint32_t)p.wrows * 128u;
      uint32_t n = 0, as = 0, aph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        if (p.w_resident && tile != (int64_t)blockIdx.x) break;
        const int cot = (int)(tile % p.n_cotiles);
        for (int c = 0; c < p.nchunks; ++c) {
          for (int j = 0; j < p.k; ++j, ++n) {
            if ((n & 1u) == mine) {
              mbar_wait(&a_empty[as], aph ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&a_full[as], a_bytes);
                tma_load_3d(a_st + as * UA_SLOT_BYTES, &tmap_w, c * 64, cot * CW, j, &a_full[as]);
              }
              __syncwarp();
            }
            if (++as == (uint32_t)p.a_stages) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: the whole warp walks the (warp-uniform) loops so
    // that descriptors and barrier addresses live in uniform registers; one elected lane issues tcgen05.mma / commit
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t desc0 = make_smem_desc(0, 128, 0);
      const uint32_t dhi = (uint32_t)(desc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)desc0 + (smem_u32(a_st) >> 4), x_lo0 = (uint32_t)desc0 + (smem_u32(x_st) >> 4);
      const uint32_t tap_step = (uint32_t)(p.dil * 128) >> 4;
      const uint32_t a_stages = (uint32_t)p.a_stages;
      const bool resident = p.w_resident != 0;
      const int k = p.k;
      uint32_t as = 0, aph = 0, xs = 0, xph = 0, acc = 0, accph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const bool w_wait = !resident || tile == (int64_t)blockIdx.x;
        if (resident) as = 0;
        mbar_wait(&t_empty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        uint32_t accum = 0;
        for (int c = 0; c < p.nchunks; ++c) {
          int nkk = (p.Cin_p - c * 64 + 15) >> 4;      // K steps that hold real channels
          nkk = nkk > 4 ? 4 : nkk;
          mbar_wait(&x_full[xs], xph);
          tc_fence_after();
          uint32_t b_lo = x_lo0 + xs * (uint32_t)(UA_X_STAGE_BYTES >> 4);
          for (int j = 0; j < k; ++j, b_lo += tap_step) {
            if (w_wait) {
              mbar_wait(&a_full[as], aph);
              tc_fence_after();
            }
            const uint32_t a_lo = a_lo0 + as * (uint32_t)(UA_SLOT_BYTES >> 4);
            if (elect_one()) {
              const uint64_t da = ((uint64_t)dhi << 32) | a_lo, db = ((uint64_t)dhi << 32) | b_lo;
              umma_f16_ss(d_tmem, da, db, idesc, accum);
              for (int kk = 1; kk < nkk; ++kk) umma_f16_ss(d_tmem, da + 2 * kk, db + 2 * kk, idesc, 1u);
              if (!resident) umma_commit(&a_empty[as]);
            }
            __syncwarp();
            accum = 1;
            if (++as == a_stages) { as = 0; aph ^= 1; }
          }
          if (elect_one()) umma_commit(&x_empty[xs]);
          __syncwarp();
          if (++xs == UA_X_STAGES) { xs = 0; xph ^= 1; }
        }
        if (elect_one()) umma_commit(&t_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps 2..9: bias + Activation1d + bf16, private TMA stores
    const int g = warp % 4;                       // TMEM lane group of this warp
    const int half = (warp - 2) >> 2;             // the two warps of a lane group
    const int lane0 = (g * 32) % p.LR;            // first channel (within the tile) held by this warp's lanes
    const int replica = (g * 32) / p.LR;
    const int ch = lane0 + lane;
    const bool warp_ok = lane0 < CW;
    const bool lane_ok = ch < CW;
    const int wbox = warp_ok ? (CW - lane0 < 32 ? CW - lane0 : 32) : 0;
    const void* omap = wbox == 32 ? (const void*)&tmap_out : (const void*)&tmap_out_tail;
    const int seg = replica * 2 + half;
    const uint32_t stg = smem_u32(o_st) + (uint32_t)(warp - 2) * 2u * UA_STAGE_BYTES;
    uint32_t acc = 0, accph = 0, nstore = 0;
    int last_cot = -1;
    float bv = 0.f, a = 1.f, ib = 1.f;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const UATile t = ua_tile(p, tile);
      if (t.cot != last_cot) {
        last_cot = t.cot;
        if (lane_ok) {
          const int c = t.cot * CW + ch;
          bv = p.bias ? __ldg(p.bias + c) : 0.f;
          a = expf(__ldg(p.alpha_log + c));
          ib = 1.0f / (expf(__ldg(p.beta_log + c)) + 1e-9f);
        }
      }
      mbar_wait(&t_full[acc], accph);
      tc_fence_after();
      const int tseg = t.t0 + seg * p.S;
      if (warp_ok && tseg < p.T) {
        const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + acc * 256u + (uint32_t)(seg * p.S);
        const bool interior = tseg >= UA_LEAD && tseg + p.S + 5 <= p.T - 1;
        const int64_t row0 = (int64_t)t.b * p.T + tseg;
        const float* rrow = RES ? p.res + (row0 - 6) * p.out_ld + (t.cot * CW + ch) : nullptr;
        float* yrow = RES ? p.y + row0 * p.out_ld + (t.cot * CW + ch) : nullptr;
        if (interior)
          ua_segment<false, RES>(p, taddr, tseg, bv, a, ib, stg, wbox, lane, lane_ok, omap, t.cot * CW + lane0, t.b, nstore,
                                 rrow, yrow);
        else
          ua_segment<true, RES>(p, taddr, tseg, bv, a, ib, stg, wbox, lane, lane_ok, omap, t.cot * CW + lane0, t.b, nstore,
                                rrow, yrow);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
    __syncwarp();
    if (elect_one()) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host side ----
static bool ua_plan(const ConvArgs& a, UAParams& p) {
  if (a.accum || a.scale != 1.f || a.bias_bs != 0) return false;
  if (a.in_dtype != BVG_BF16 || a.w_dtype != BVG_BF16 || a.out_dtype != BVG_BF16) return false;
  if (a.Cin_p % 8 != 0 || a.Cout_r % 128 != 0 || a.Cout_n <= 0 || a.Cout_n % 8 != 0) return false;
  if (a.T <= 0 || a.T > 0x3fffffffLL || a.B <= 0) return false;
  const int halo = (a.k - 1) * a.dil;
  const int ncot = (int)ceil_div(a.Cout_n, 128);
  if (a.Cout_n % ncot) return false;
  const int CW = a.Cout_n / ncot;
  if (CW % 8) return false;
  if (((int64_t)a.out_ld * 2) % 16 || ((int64_t)a.Cin_p * 2) % 16) return false;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.out);
  if (al & 15) return false;
  p.bias = a.bias;
  p.res = a.res; p.y = nullptr; p.out_ld = a.out_ld;
  if (a.res && (reinterpret_cast<uintptr_t>(a.res) & 3)) return false;
  p.B = a.B; p.T = (int)a.T;
  p.Cin_p = a.Cin_p; p.nchunks = (int)ceil_div(a.Cin_p, 64);
  p.k = a.k; p.dil = a.dil; p.center = (a.k - 1) / 2;
  p.CW = CW; p.n_cotiles = ncot;
  p.rep = CW <= 32 ? 4 : (CW <= 64 ? 2 : 1);
  p.LR = 128 / p.rep;
  if (p.rep > 1 && weight_replica_rows(a.Cout_n, a.Cout_r) != p.LR) return false;
  p.wrows = p.rep == 1 ? round_up(CW, 8) : 128;
  if ((int64_t)(ncot - 1) * CW + p.wrows > a.Cout_r) return false;
  p.a_stages = UA_A_SLOTS;
  p.w_resident = (ncot == 1 && a.k * p.nchunks <= p.a_stages) ? 1 : 0;
  p.S = p.rep == 1 ? 120 : (p.rep == 2 ? 60 : 24);
  p.Rs = p.rep == 4 ? 12 : 30;
  // small problems (one short utterance): if the half-width tiles (N = 144 instead of 256: ~100 instead of ~150 cycles per
  // tcgen05.mma, tools/umma_probe.cu) still fit one wave, every layer finishes sooner - a 768-channel k = 11 layer on one 2 s
  // utterance is 18 tiles of 528 MMAs otherwise
  if (p.rep == 1 && (int64_t)a.B * ceil_div(a.T, 120) * ncot <= umma_sm_count()) p.S = 60;
  p.NOUT = 2 * p.rep * p.S;
  p.NT = round_up(p.NOUT + 11, 16);
  if (p.NT + halo > UA_MAX_X_ROWS - 8) return false;
  p.x_nbox = (p.NT + halo) <= 256 ? 1 : 2;
  p.x_box_rows = round_up((p.NT + halo + p.x_nbox - 1) / p.x_nbox, 8);
  if (p.x_nbox * p.x_box_rows > UA_MAX_X_ROWS) return false;
  p.n_ttiles = (int)ceil_div(a.T, p.NOUT);
  p.n_tiles = (int64_t)a.B * p.n_ttiles * ncot;
  return true;
}

bool conv_umma2a_supported(const ConvArgs& a) {
  UAParams p;
  return ua_plan(a, p);
}

int conv_umma2a_launch(const ConvArgs& a, const float* alpha_log, const float* beta_log, const Taps& taps,
                       cudaStream_t st, float* y_out) {
  if (a.B <= 0 || a.T <= 0) return BVG_OK;
  UAParams p;
  if (!ua_plan(a, p)) BVG_FAIL(BVG_EINVAL, "conv_umma2a: unsupported layer shape/dtype");
  if (!alpha_log || !beta_log) BVG_FAIL(BVG_EINVAL, "conv_umma2a: null activation parameters");
  if ((a.res != nullptr) != (y_out != nullptr)) BVG_FAIL(BVG_EINVAL, "conv_umma2a: residual and y output go together");
  if (a.res && static_cast<const void*>(a.res) == static_cast<const void*>(y_out))
    BVG_FAIL(BVG_EINVAL, "conv_umma2a: y must not alias the residual (tiles read each other's halo rows)");
  p.y = y_out;
  p.alpha_log = alpha_log; p.beta_log = beta_log;
  make_taps_packed(&p.tp, taps);
  CUtensorMap mx, mw, mo, mt;
  int rc = make_map_any(&mx, a.in, 2, (uint64_t)a.Cin_p, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.Cin_p, 64,
                        (uint32_t)p.x_box_rows, 1, 128);
  if (rc) return rc;
  rc = make_map_any(&mw, a.w, 2, (uint64_t)a.Cin_p, (uint64_t)a.Cout_r, (uint64_t)a.k, (uint64_t)a.Cin_p, 64,
                    (uint32_t)p.wrows, 1, 128);
  if (rc) return rc;
  // output boxes: 32 channels x Rs rows per warp and sub-segment (and the narrower last box of tiles with CW % 32 != 0)
  const int tailw = p.CW % 32;
  const int fullw = p.CW >= 32 ? 32 : tailw;
  rc = make_map_any(&mo, a.out, 2, (uint64_t)a.Cout_n, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.out_ld,
                    (uint32_t)fullw, (uint32_t)p.Rs, 1, 0);
  if (rc) return rc;
  mt = mo;
  if (tailw && p.CW >= 32) {
    rc = make_map_any(&mt, a.out, 2, (uint64_t)a.Cout_n, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.out_ld,
                      (uint32_t)tailw, (uint32_t)p.Rs, 1, 0);
    if (rc) return rc;
  }
  const int sms = umma_sm_count();
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  const int smem = a.own_sm ? UA_SMEM_BYTES : UA_SMEM_USED;
  if (a.res) {
    static std::atomic<unsigned long long> attr_done_res{0};
    if (int rc_ = smem_attr_once(conv_umma2a_kernel<true>, UA_SMEM_BYTES, attr_done_res)) return rc_;
    launch_pdl(conv_umma2a_kernel<true>, grid, UA_THREADS, smem, st, mx, mw, mo, mt, p);
  } else {
    static std::atomic<unsigned long long> attr_done_plain{0};
    if (int rc_ = smem_attr_once(conv_umma2a_kernel<false>, UA_SMEM_BYTES, attr_done_plain)) return rc_;
    launch_pdl(conv_umma2a_kernel<false>, grid, UA_THREADS, smem, st, mx, mw, mo, mt, p);
  }
  BVG_LAUNCHED();
  return BVG_OK;
}

bool conv_act_fused_supported(const ConvArgs& a) { return conv_umma2a_supported(a); }
int conv_act_fused_launch(const ConvArgs& a, const float* alpha_log, const float* beta_log, const Taps& taps,
                          cudaStream_t st, float* y_out) {
  return conv_umma2a_launch(a, alpha_log, beta_log, taps, st, y_out);
}

}  // namespace bvg
