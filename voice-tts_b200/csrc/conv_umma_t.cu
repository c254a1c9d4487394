// Time-major tcgen05 implicit-GEMM conv for layers with few output channels
// (Cout_n <= 128: the 96/48/24-channel stages and their ConvTranspose1d's).
//
//   D[t, co] = sum_tap sum_ci  X[t + (tap-center)*dil][ci] * Wp[tap][co][ci]
//
//   A operand (M = 128 time rows, K-major): rows of the staged activation tile; tap j of sub-tile m reads
//                                           rows [m*128 + j*dil, +128) of the SAME shared-memory tile
//   B operand (N = Cout_n channels, K-major): weight tile Wp[tap][0:N][ci0:ci0+KC]  (one TMA box per tap/chunk,
//                                           shared by all NSUB sub-tiles of the CTA's time tile)
//   D (fp32, TMEM): NSUB accumulators of 128 lanes (time) x N columns, double buffered
//
// Compared with the channel-major kernel (conv_umma.cu) at these widths: no MMA rows are
// spent on channel padding (M is time), every epilogue warp owns 32 useful TMEM lanes (time
// rows) instead of one warp owning all valid channels, and a lane holds 32 consecutive
// channels of ONE output row, so the [time][channel] tile is transposed inside each warp
// (swizzled 4 KB scratch, __syncwarp only) into fully coalesced 512-byte global accesses.
// The residual / accumulate operands of the NEXT tile are prefetched into L2 while the
// current tile is stored.
//
// Warp roles as in conv_umma.cu: warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer,
// warps 2-5 epilogue.  reference semantics: torch Conv1d/ConvTranspose1d (bigvgan.py:59-66,76-83,306-312).
#include "umma_common.cuh"

namespace bvg {

constexpr int UT_EPI_WARPS = 8;                       // two epilogue warps per TMEM lane group
constexpr int UT_THREADS = 64 + 32 * UT_EPI_WARPS;
constexpr int UT_W_STAGES = 4;                        // streaming mode: ring of tap groups
constexpr int UT_X_STAGES_MAX = 4;                    // activation-tile ring depth (2..4, by smem budget)
constexpr int UT_X_STAGE_MAX = 48 * 1024;             // activation tile budget per stage
constexpr int UT_SCRATCH_BYTES = UT_EPI_WARPS * 32 * 32 * 4;   // per-warp 32x32 fp32 transpose scratch
constexpr int UT_SMEM_MAX = 226 * 1024;
constexpr int UT_W_RESIDENT_MAX = 110 * 1024;         // weights of a whole layer stay in smem below this

struct UmmaTParams {
  const float* bias;
  void* out;
  const float* res;
  const float* accum;
  float scale;
  int out_bf16;
  int B, T;
  int Cin_p, N, Np, out_ld;   // N = Cout_n (multiple of 16), Np = TMEM columns per accumulator
  int k, dil, center;
  int KC, nchunks;
  int NSUB, MT;               // sub-tiles of 128 rows per tile, MT = NSUB*128
  int x_box_rows, x_nbox;     // TMA boxes of the activation tile
  int n_ttiles;
  int64_t n_tiles;            // B * n_ttiles
  int base_mode;
  int w_resident;             // 1: all k*nchunks weight tiles loaded once and kept in shared memory
  int tps;                    // streaming mode: taps per ring stage (one TMA box {KC, N, tps})
  int wtile_bytes;            // N * KC * 2
  int w_region_bytes, x_stage_bytes;
  int x_stages;               // activation-tile ring depth
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <bool RES, bool ACC, bool BF16OUT>
__global__ void __launch_bounds__(UT_THREADS, 1)
conv_umma_t_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const UmmaTParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_st = smem;
  unsigned char* x_st = smem + p.w_region_bytes;
  float* scratch = reinterpret_cast<float*>(x_st + p.x_stages * p.x_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + UT_SCRATCH_BYTES);
  const int w_stage_bytes = p.w_resident ? 0 : p.w_region_bytes / UT_W_STAGES;
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + UT_W_STAGES;
  uint64_t* x_full = w_empty + UT_W_STAGES;
  uint64_t* x_empty = x_full + UT_X_STAGES_MAX;
  uint64_t* t_full = x_empty + UT_X_STAGES_MAX;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int row_bytes = p.KC * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < UT_W_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < UT_X_STAGES_MAX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], UT_EPI_WARPS); }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t xbox_bytes = (uint32_t)p.x_box_rows * row_bytes;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t ws = 0, wph = 0, xs = 0, xph = 0;
      if (p.w_resident) {
        // whole layer: k*nchunks tiles of [N x KC], one mbarrier (w_full[0]) for all of them
        mbar_expect_tx(&w_full[0], (uint32_t)(p.k * p.nchunks) * (uint32_t)p.wtile_bytes);
        for (int c = 0; c < p.nchunks; ++c)
          for (int j = 0; j < p.k; ++j)
            tma_load_3d(w_st + (c * p.k + j) * p.wtile_bytes, &tmap_w, c * p.KC, 0, j, &w_full[0]);
      }
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int tt = (int)(tile % p.n_ttiles);
        const int b = (int)(tile / p.n_ttiles);
        const int t0 = tt * p.MT;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&x_empty[xs], xph ^ 1);
          mbar_expect_tx(&x_full[xs], (uint32_t)p.x_nbox * xbox_bytes);
          unsigned char* dstx = x_st + xs * p.x_stage_bytes;
          const int trow = t0 - p.center * p.dil;
          for (int bx = 0; bx < p.x_nbox; ++bx)
            tma_load_3d(dstx + bx * xbox_bytes, &tmap_x, c * p.KC, trow + bx * p.x_box_rows, b, &x_full[xs]);
          if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
          if (!p.w_resident) {
            for (int j0 = 0; j0 < p.k; j0 += p.tps) {
              // the box always has tps taps; taps beyond k are zero-filled by TMA (and never used)
              mbar_wait(&w_empty[ws], wph ^ 1);
              mbar_expect_tx(&w_full[ws], (uint32_t)p.tps * (uint32_t)p.wtile_bytes);
              tma_load_3d(w_st + ws * w_stage_bytes, &tmap_w, c * p.KC, 0, j0, &w_full[ws]);
              if (++ws == UT_W_STAGES) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: the whole warp walks the loop with
    // warp-uniform values (descriptors stay in uniform registers); one elected lane issues.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // per MMA only the 14-bit start-address field (addr >> 4) of the descriptor template advances
    const uint64_t desc0 = make_smem_desc(0, row_bytes, 0);
    const uint32_t sub_step = (uint32_t)(128 * row_bytes) >> 4;      // next 128-row sub-tile
    const uint32_t tap_step = (uint32_t)(p.dil * row_bytes) >> 4;    // next tap: dil rows further
    const int nkk = p.KC / 16;
    const uint32_t x_base = smem_u32(x_st), w_base = smem_u32(w_st);
    uint32_t ws = 0, wph = 0, xs = 0, xph = 0, acc = 0, accph = 0;
    if (p.w_resident) {
      mbar_wait(&w_full[0], 0);
      tc_fence_after();
    }
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&t_empty[acc], accph ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + acc * (uint32_t)(p.NSUB * p.Np);
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(&x_full[xs], xph);
        tc_fence_after();
        const uint32_t x_addr = x_base + xs * p.x_stage_bytes;
        for (int j0 = 0; j0 < p.k; j0 += p.tps) {
          uint32_t wg_addr;
          if (p.w_resident) {
            wg_addr = w_base + (c * p.k + j0) * p.wtile_bytes;
          } else {
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            wg_addr = w_base + ws * w_stage_bytes;
          }
          const int jend = (j0 + p.tps < p.k) ? j0 + p.tps : p.k;
          if (elect_one()) {
            for (int j = j0; j < jend; ++j) {
              const uint64_t db0 = desc0 + ((wg_addr + (uint32_t)(j - j0) * p.wtile_bytes) >> 4);
              uint64_t da = desc0 + (x_addr >> 4) + (uint32_t)j * tap_step;
              uint32_t d_col = d_base;
              const uint32_t accumulate = (c | j) ? 1u : 0u;
              for (int m = 0; m < p.NSUB; ++m) {
                umma_f16_ss(d_col, da, db0, idesc, accumulate);
                for (int kk = 1; kk < nkk; ++kk) umma_f16_ss(d_col, da + 2 * kk, db0 + 2 * kk, idesc, 1u);
                da += sub_step;
                d_col += p.Np;
              }
            }
            if (!p.w_resident) umma_commit(&w_empty[ws]);
          }
          __syncwarp();
          if (!p.w_resident) {
            if (++ws == UT_W_STAGES) { ws = 0; wph ^= 1; }
          }
        }
        if (elect_one()) umma_commit(&x_empty[xs]);
        __syncwarp();
        if (++xs == (uint32_t)p.x_stages) { xs = 0; xph ^= 1; }
      }
      if (elect_one()) umma_commit(&t_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  } else {
    // -------------------------------------------------- epilogue warps 2..9
    // Two warps per TMEM lane group; the (sub-tile, 32-channel chunk) work items of a tile
    // alternate between them.  RES / ACC / BF16OUT are compile-time so the per-vector code has
    // no branches; full 32-row blocks skip the row predicates.
    const int g = warp % 4;                 // TMEM lane group (time rows 32g..32g+31 of each sub-tile)
    const int ew = warp - 2;                // 0..7
    const int half = ew >> 2;               // which of the two warps of this lane group
    const int etid = ew * 32 + lane;
    float* my = scratch + ew * 1024;        // this warp's 32x32 transpose scratch
    const int r8 = lane >> 3, c8 = lane & 7; // transposed role: row 4q + r8, channels 4*c8 .. 4*c8+3 of the chunk
    const int row_stride = 4 * p.out_ld;    // elements between the rows of consecutive q
    const int nchunk32 = (p.N + 31) / 32;
    uint32_t acc = 0, accph = 0;

    // L2 prefetch of the residual / accumulate rows of one tile (fire and forget)
    auto prefetch_tile = [&](int64_t tile) {
      if (!RES && !ACC) return;
      const int tt = (int)(tile % p.n_ttiles);
      const int b = (int)(tile / p.n_ttiles);
      const int t0 = tt * p.MT;
      int rows = p.T - t0;
      if (rows > p.MT) rows = p.MT;
      const int64_t base = ((int64_t)b * p.T + t0) * p.out_ld;
      const int64_t nbytes = (int64_t)rows * p.out_ld * 4;   // rows are contiguous: out_ld == N for these layers
      for (int64_t o = (int64_t)etid * 128; o < nbytes; o += 128 * 32 * UT_EPI_WARPS) {
        if (RES) prefetch_l2(reinterpret_cast<const char*>(p.res + base) + o);
        if (ACC) prefetch_l2(reinterpret_cast<const char*>(p.accum + base) + o);
      }
    };
    if ((int64_t)blockIdx.x < p.n_tiles) prefetch_tile(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int tt = (int)(tile % p.n_ttiles);
      const int b = (int)(tile / p.n_ttiles);
      const int t0 = tt * p.MT;
      if (tile + gridDim.x < p.n_tiles) prefetch_tile(tile + gridDim.x);

      mbar_wait(&t_full[acc], accph);
      tc_fence_after();
      const uint32_t d_base = tmem_base + ((uint32_t)(g * 32) << 16) + acc * (uint32_t)(p.NSUB * p.Np);
      const int nitems = p.NSUB * nchunk32;
      for (int w = half; w < nitems; w += 2) {
        const int m = w / nchunk32, cc = w - m * nchunk32;
        const int trow0 = t0 + m * 128 + g * 32;            // first time row of this warp's 32-row block
        const int nrows = p.T - trow0;                       // valid rows from trow0 on (may be <= 0)
        const int cvalid = p.N - cc * 32;                    // >= 16
        const bool lane_ok_c = c8 * 4 < cvalid;
        const bool full = nrows >= 32 && cvalid >= 32;
        const int64_t e0 = ((int64_t)b * p.T + trow0 + r8) * p.out_ld + cc * 32 + c8 * 4;
        const float* resp = RES ? p.res + e0 : nullptr;
        const float* accp = ACC ? p.accum + e0 : nullptr;
        // residual / accumulate operands first: their latency overlaps the TMEM load + transpose
        float4 rv[8], av[8];
        if (RES || ACC) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bool ok = full || (lane_ok_c && (4 * q + r8) < nrows);
            if (RES) rv[q] = ok ? *reinterpret_cast<const float4*>(resp + q * row_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ACC) av[q] = ok ? *reinterpret_cast<const float4*>(accp + q * row_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && lane_ok_c) bv = __ldg(reinterpret_cast<const float4*>(p.bias + cc * 32 + c8 * 4));

        uint32_t v[32];
        tmem_ld_32x32(d_base + m * p.Np + cc * 32, v);   // lane = time row, v[i] = channel cc*32+i
        tmem_ld_wait();
        // transpose through the warp's scratch: row = lane, 16-byte chunk i stored at chunk (i ^ (lane & 7))
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 f = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                 __uint_as_float(v[4 * i + 3]));
          *reinterpret_cast<float4*>(my + lane * 32 + ((i ^ (lane & 7)) << 2)) = f;
        }
        __syncwarp();
        const float sc = p.scale;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int rr = 4 * q + r8;
          const float4 d = *reinterpret_cast<const float4*>(my + rr * 32 + ((c8 ^ (rr & 7)) << 2));
          float4 y;
          y.x = d.x + bv.x; y.y = d.y + bv.y; y.z = d.z + bv.z; y.w = d.w + bv.w;
          if (RES) { y.x += rv[q].x; y.y += rv[q].y; y.z += rv[q].z; y.w += rv[q].w; }
          y.x *= sc; y.y *= sc; y.z *= sc; y.w *= sc;
          if (ACC) { y.x += av[q].x; y.y += av[q].y; y.z += av[q].z; y.w += av[q].w; }
          const bool ok = full || (lane_ok_c && rr < nrows);
          if (ok) {
            if (BF16OUT) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + e0 + q * row_stride) = pk;
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + e0 + q * row_stride) = y;
            }
          }
        }
      }
      // all TMEM reads of this tile are complete (tcgen05.wait::ld above): release the accumulators
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// shared-memory / TMEM plan of one layer; false if the layer does not fit this kernel
static bool plan_t(const ConvArgs& a, int variant, UmmaTParams& p, int* smem_out) {
  p.bias = a.bias; p.out = a.out; p.res = a.res; p.accum = a.accum; p.scale = a.scale;
  p.out_bf16 = a.out_dtype == BVG_BF16;
  p.B = a.B; p.T = (int)a.T; p.Cin_p = a.Cin_p; p.N = a.Cout_n; p.out_ld = a.out_ld;
  p.Np = round_up(a.Cout_n, 32);
  p.k = a.k; p.dil = a.dil; p.center = (a.k - 1) / 2;
  p.KC = (a.Cin_p % 64 == 0) ? 64 : (a.Cin_p % 32 == 0 ? 32 : 16);
  p.nchunks = a.Cin_p / p.KC;
  const int row_bytes = p.KC * 2;
  const int halo = (a.k - 1) * a.dil;
  // sub-tiles per CTA tile: bounded by TMEM (2 buffers x NSUB x Np <= 512 columns), by the
  // activation-stage budget, and by the amount of work (keep >= ~2 tiles per SM)
  p.wtile_bytes = p.N * row_bytes;
  const int w_total = p.k * p.nchunks * p.wtile_bytes;
  p.w_resident = w_total <= UT_W_RESIDENT_MAX ? 1 : 0;
  if (p.w_resident) {
    p.tps = 1;
    p.w_region_bytes = round_up(w_total, 1024);
  } else {
    // stream groups of taps: ~16 KB per ring stage
    p.tps = 16384 / p.wtile_bytes;
    if (p.tps < 1) p.tps = 1;
    if (p.tps > p.k) p.tps = p.k;
    p.w_region_bytes = UT_W_STAGES * round_up(p.tps * p.wtile_bytes, 1024);
  }
  const int x_total_budget = UT_SMEM_MAX - 1024 - 256 - UT_SCRATCH_BYTES - p.w_region_bytes;
  const int x_budget = x_total_budget / 2;
  const int x_cap = x_budget < UT_X_STAGE_MAX ? x_budget : UT_X_STAGE_MAX;
  int nsub = 256 / p.Np;
  if (nsub > 4) nsub = 4;
  while (nsub > 1 && round_up(nsub * 128 + halo + 24, 8) * row_bytes > x_cap) --nsub;
  const int sms = umma_sm_count();
  while (nsub > 1 && (int64_t)a.B * ceil_div(a.T, nsub * 128) < 2LL * sms) --nsub;
  p.NSUB = nsub;
  p.MT = nsub * 128;
  const int xrows = p.MT + halo;
  p.x_nbox = (int)ceil_div(xrows, 256);
  p.x_box_rows = round_up((int)ceil_div(xrows, p.x_nbox), 8);
  p.x_stage_bytes = round_up(p.x_nbox * p.x_box_rows * row_bytes, 1024);
  if (p.x_stage_bytes > x_cap + 1024 || p.x_stage_bytes * 2 + p.w_region_bytes + UT_SCRATCH_BYTES + 1280 > UT_SMEM_MAX)
    return false;
  // deeper activation ring when it fits: the per-tile MMA work of these layers is short, so the
  // TMA latency of the (narrow-row) activation tile must be hidden by running several tiles ahead
  p.x_stages = x_total_budget / p.x_stage_bytes;
  if (p.x_stages > UT_X_STAGES_MAX) p.x_stages = UT_X_STAGES_MAX;
  if (p.x_stages < 2) p.x_stages = 2;
  p.n_ttiles = (int)ceil_div(a.T, p.MT);
  p.n_tiles = (int64_t)a.B * p.n_ttiles;
  p.base_mode = variant & 1;
  if (a.out_ld != a.Cout_n || a.Cout_n > 128 || a.Cout_n % 16) return false;
  *smem_out = 1024 + p.w_region_bytes + p.x_stages * p.x_stage_bytes + UT_SCRATCH_BYTES + 256;
  return true;
}

bool conv_umma_t_fits(const ConvArgs& a) {
  UmmaTParams p;
  int smem = 0;
  return plan_t(a, 0, p, &smem);
}

int conv_umma_t_launch(const ConvArgs& a, int variant, cudaStream_t st) {
  UmmaTParams p;
  int smem_bytes = 0;
  if (!plan_t(a, variant, p, &smem_bytes))
    BVG_FAIL(BVG_EINVAL, "conv_umma_t: layer does not fit (Cin_p=%d N=%d k=%d dil=%d)", a.Cin_p, a.Cout_n, a.k, a.dil);
  const int row_bytes = p.KC * 2;
  const int sms = umma_sm_count();

  CUtensorMap mx, mw;
  int rc = make_map_3d(&mx, a.in, (uint64_t)a.Cin_p, (uint64_t)a.T, (uint64_t)a.B, (uint32_t)p.KC,
                       (uint32_t)p.x_box_rows, row_bytes);
  if (rc) return rc;
  rc = make_map_4d_w(&mw, a.w, (uint64_t)a.Cin_p, (uint64_t)a.Cout_r, (uint64_t)a.k, (uint32_t)p.KC, (uint32_t)p.N,
                     (uint32_t)p.tps, row_bytes);
  if (rc) return rc;
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
#define BVG_T_LAUNCH(R, A, O)                                                                                         \
  do {                                                                                                                \
    BVG_CUDA(cudaFuncSetAttribute(conv_umma_t_kernel<R, A, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, UT_SMEM_MAX)); \
    conv_umma_t_kernel<R, A, O><<<grid, UT_THREADS, smem_bytes, st>>>(mx, mw, p);                                     \
  } while (0)
  const int sel = (p.res ? 4 : 0) | (p.accum ? 2 : 0) | (p.out_bf16 ? 1 : 0);
  switch (sel) {
    case 0: BVG_T_LAUNCH(false, false, false); break;
    case 1: BVG_T_LAUNCH(false, false, true); break;
    case 2: BVG_T_LAUNCH(false, true, false); break;
    case 3: BVG_T_LAUNCH(false, true, true); break;
    case 4: BVG_T_LAUNCH(true, false, false); break;
    case 5: BVG_T_LAUNCH(true, false, true); break;
    case 6: BVG_T_LAUNCH(true, true, false); break;
    default: BVG_T_LAUNCH(true, true, true); break;
  }
#undef BVG_T_LAUNCH
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
