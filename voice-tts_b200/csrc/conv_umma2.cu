// Dilated Conv1d / polyphase ConvTranspose1d as an implicit GEMM on tcgen05 - second generation
// of the channel-major kernel (conv_umma.cu), rebuilt around measurements taken on B200
// (tools/umma_probe.cu, tools/tma_probe.cu, profiles/r01_umma_probe.txt, profiles/r01_tma_probe.txt):
//   * an M=128, K=16 tcgen05.mma never takes less than ~90 cycles (A-operand fetch), whatever N is, so
//     narrow layers put TIME on N (up to 256 columns per instruction) and channels on M;
//   * operand rows narrower than 128 bytes (SWIZZLE_64B / 32B) cost 136 / 173 cycles per MMA instead
//     of 90-129: operands are always staged as 128-byte rows - layers with fewer than 64 input
//     channels let TMA zero-fill the row tail and skip the all-zero K steps;
//   * ONE thread gets at most one TMA operation through per ~400 cycles, whatever its size (two issuing
//     warps get two): TMA work is therefore spread over dedicated issuer warps, boxes are as large as
//     possible, and no thread that computes ever issues or waits for a copy;
//   * a TMA store drains shared memory at ~29 B/clk/SM.
//
//   D[co, t] = sum_tap sum_ci  Wp[tap][co][ci] * X[t + (tap-center)*dil][ci]
//
//   A (M = 128 rows, K-major, 128-byte rows): weight rows of one out-channel tile (CW <= 128 channels).
//       Layers with <= 64 (<= 32) output channels store 2 (4) copies of their rows in the 128-row
//       tile at pack time (layout.cu: weight_replica_rows), so that every TMEM lane group - hence
//       every epilogue warp - holds a copy of the result and takes a share of the time columns.
//       Weights stay resident in shared memory when all taps fit the ring, else they stream.
//   B (N = NT time rows, K-major): channels-last activation tile incl. halo, loaded once per 64-channel
//       chunk and reused by all taps through row-shifted descriptors; TMA zero fill = conv padding.
//   D (fp32, TMEM): 128 lanes x NT columns, double buffered.
//
// Warps (14):  0 activation-tile loads | 1 TMEM alloc + MMA issue | 2-9 epilogue math |
//              10,11 weight loads (alternate ring stages) | 12 output stores | 13 residual/accumulate loads.
// The epilogue works in blocks of CB time rows: a math warp reads 16 (8) TMEM columns of its 32 lanes,
// adds bias and the residual / accumulate rows that warp 13 has already brought into shared memory
// ([time][channel] fp32), writes the result into a [time][channel] staging block (conflict-free: lanes =
// consecutive channels) and arrives on an mbarrier; warp 12 hands the finished block to a TMA tensor
// store (rows beyond T are clipped by the tensor map).  All hand-offs are mbarriers - there is no
// CTA-wide barrier in the steady state and no per-thread global memory access.
//
// reference semantics: torch Conv1d/ConvTranspose1d as built in bigvgan.py:59-66,76-83,285-287,306-312.
#include "umma_common.cuh"

namespace bvg {

constexpr int U2_EPI_WARPS = 8;
constexpr int U2_WARPS = 14;
constexpr int U2_THREADS = 32 * U2_WARPS;
constexpr int U2_SLOT_BYTES = 16384;                 // one weight stage (128 rows x 128 B) or one staging block
constexpr int U2_AIN_SLOTS = 7;                      // weight ring (+ 3 residual staging slots when used)
constexpr int U2_IN_SLOTS = 3;
constexpr int U2_X_STAGES = 2;
constexpr int U2_MAX_X_ROWS = 320;
constexpr int U2_X_STAGE_BYTES = U2_MAX_X_ROWS * 128;
constexpr int U2_OUT_SLOTS = 2;
constexpr int U2_SMEM_USED = 1024 + U2_AIN_SLOTS * U2_SLOT_BYTES + U2_X_STAGES * U2_X_STAGE_BYTES +
                             U2_OUT_SLOTS * U2_SLOT_BYTES + 512;
// The launch asks for the whole 227 KB opt-in maximum: with the 1 KB the system reserves per CTA that is the SM's entire
// 228 KB carve-out, so no CTA of any other kernel (not even one without shared memory) can become resident beside this
// persistent CTA.  Co-resident blocks of another stream were the condition under which older revisions returned corrupted
// tiles (root cause: hand-off arrives overtaking unfinished loads, fixed by loads_landed() - DESIGN.md section 7.1); owning
// the SM costs nothing and stays as a second line of defence.
constexpr int U2_SMEM_BYTES = 227 * 1024;
static_assert(U2_SMEM_USED <= U2_SMEM_BYTES, "shared-memory plan exceeds the opt-in maximum");

struct U2Params {
  const float* bias;
  int64_t bias_bs;   // per-utterance bias stride (0: shared)
  float scale;
  int B, T;
  int Cin_p, nchunks;
  int k, dil, center;
  int NT, x_box_rows, x_nbox;
  int n_ttiles, n_cotiles;
  int64_t n_tiles;
  int CW;            // out channels per tile
  int rep, LR;       // weight replicas in the 128 MMA rows, lanes per replica (128 / rep)
  int wrows;         // weight rows per TMA box
  int a_stages;
  int w_resident;    // all k*nchunks weight tiles are loaded once and stay in the ring slots
  int CB;            // time rows per epilogue block = NCOL * 2 * rep
};

struct U2Tile {
  int cot, b, t0, nb_end;
};
__device__ __forceinline__ U2Tile u2_tile(const U2Params& p, int64_t tile) {
  U2Tile t;
  t.cot = (int)(tile % p.n_cotiles);
  const int64_t r = tile / p.n_cotiles;
  const int tt = (int)(r % p.n_ttiles);
  t.b = (int)(r / p.n_ttiles);
  t.t0 = tt * p.NT;
  t.nb_end = p.T - t.t0;
  if (t.nb_end > p.NT) t.nb_end = p.NT;
  return t;
}

// NIN: 0 = plain, 1 = + residual, 2 = + residual and accumulate operand
template <int NIN, bool BF16OUT>
__global__ void __launch_bounds__(U2_THREADS, 1)
conv_umma2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                  const __grid_constant__ CUtensorMap tmap_acc, const U2Params p) {
  constexpr int NCOL = NIN == 2 ? 8 : 16;            // TMEM columns per warp per block
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_st = smem;
  unsigned char* in_st = smem + (U2_AIN_SLOTS - U2_IN_SLOTS) * U2_SLOT_BYTES;
  unsigned char* x_st = smem + U2_AIN_SLOTS * U2_SLOT_BYTES;
  unsigned char* out_st = x_st + U2_X_STAGES * U2_X_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_st + U2_OUT_SLOTS * U2_SLOT_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + U2_AIN_SLOTS;
  uint64_t* x_full = a_empty + U2_AIN_SLOTS;
  uint64_t* x_empty = x_full + U2_X_STAGES;
  uint64_t* t_full = x_empty + U2_X_STAGES;
  uint64_t* t_empty = t_full + 2;
  uint64_t* in_full = t_empty + 2;
  uint64_t* in_free = in_full + U2_IN_SLOTS;
  uint64_t* out_ready = in_free + U2_IN_SLOTS;
  uint64_t* out_free = out_ready + U2_OUT_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_free + U2_OUT_SLOTS);
  const uint32_t sink_u32 = smem_u32(tmem_slot + 4 + (threadIdx.x / 32));   // one sink word per warp (loads_landed, common.cuh)

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < U2_AIN_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < U2_X_STAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], U2_EPI_WARPS); }
    for (int i = 0; i < U2_IN_SLOTS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_free[i], U2_EPI_WARPS); }
    for (int i = 0; i < U2_OUT_SLOTS; ++i) { mbar_init(&out_ready[i], U2_EPI_WARPS); mbar_init(&out_free[i], 1); }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    if (NIN >= 1) tma_prefetch_desc(&tmap_res);
    if (NIN == 2) tma_prefetch_desc(&tmap_acc);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // everything above overlapped the predecessor's tail; no global access before this line (common.cuh, PDL)
  const uint32_t tmem_base = *tmem_slot;
  const int CB = p.CB, CW = p.CW;

  if (warp == 0) {
    {
      // ------------------------------------------------ activation tiles (whole warp in the loop, one elected lane issues)
      const uint32_t xbox_bytes = (uint32_t)p.x_box_rows * 128u;
      uint32_t xs = 0, xph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const U2Tile t = u2_tile(p, tile);
        const int trow = t.t0 - p.center * p.dil;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&x_empty[xs], xph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&x_full[xs], (uint32_t)p.x_nbox * xbox_bytes);
            unsigned char* dstx = x_st + xs * U2_X_STAGE_BYTES;
            for (int q = 0; q < p.x_nbox; ++q)
              tma_load_3d(dstx + q * xbox_bytes, &tmap_x, c * 64, trow + q * p.x_box_rows, t.b, &x_full[xs]);
          }
          __syncwarp();
          if (++xs == U2_X_STAGES) { xs = 0; xph ^= 1; }
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
                tma_load_3d(a_st + as * U2_SLOT_BYTES, &tmap_w, c * 64, cot * CW, j, &a_full[as]);
              }
              __syncwarp();
            }
            if (++as == (uint32_t)p.a_stages) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: the whole warp walks the (warp-uniform) loops so that
    // descriptors, barrier addresses and ring state live in uniform registers, and one elected lane issues
    // tcgen05.mma / tcgen05.commit.  (A single thread behind `if (lane == 0)` makes the compiler wrap every
    // uniform-datapath instruction in an R2UR + vote retry loop; measured on the 384/768-channel layers that
    // costs ~130 cycles per tap - 1.30 -> 1.47 PFLOP/s on the 768-channel k = 11 layer once removed.)
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
          uint32_t b_lo = x_lo0 + xs * (uint32_t)(U2_X_STAGE_BYTES >> 4);
          for (int j = 0; j < k; ++j, b_lo += tap_step) {
            if (w_wait) {
              mbar_wait(&a_full[as], aph);
              tc_fence_after();
            }
            const uint32_t a_lo = a_lo0 + as * (uint32_t)(U2_SLOT_BYTES >> 4);
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
          if (++xs == U2_X_STAGES) { xs = 0; xph ^= 1; }
        }
        if (elect_one()) umma_commit(&t_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
  } else if (warp == 12) {
    {
      // ------------------------------------------------ output stores: block cb leaves from staging buffer cb & 1
      // (one elected lane issues; bulk async-groups are per thread and elect.sync is deterministic for a given
      //  member mask, so the same lane commits and waits every time)
      uint32_t cb = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const U2Tile t = u2_tile(p, tile);
        for (int nb = 0; nb < t.nb_end; nb += CB, ++cb) {
          const uint32_t ob = cb & 1u;
          if (cb >= 1) {                       // hand the previous block's buffer back as soon as it has been read
            if (elect_one()) {
              bulk_wait_group_read<0>();
              mbar_arrive(&out_free[ob ^ 1u]);
            }
            __syncwarp();
          }
          mbar_wait(&out_ready[ob], (cb >> 1) & 1u);
          if (elect_one()) {
            tma_store_3d(&tmap_out, out_st + ob * U2_SLOT_BYTES, t.cot * CW, t.t0 + nb, t.b);
            bulk_commit_group();
          }
          __syncwarp();
        }
      }
      __syncwarp();
      if (elect_one()) bulk_wait_group<0>();
    }
  } else if (warp == 13) {
    if (NIN >= 1) {
      // ------------------------------------------------ residual / accumulate rows, 3 blocks ahead of the math warps
      const uint32_t in_bytes = (uint32_t)(CB * CW * 4 * NIN);
      uint32_t cb = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const U2Tile t = u2_tile(p, tile);
        for (int nb = 0; nb < t.nb_end; nb += CB, ++cb) {
          const uint32_t slot = cb % U2_IN_SLOTS;
          mbar_wait(&in_free[slot], ((cb / U2_IN_SLOTS) & 1u) ^ 1u);
          if (elect_one()) {
            unsigned char* dst = in_st + slot * U2_SLOT_BYTES;
            mbar_expect_tx(&in_full[slot], in_bytes);
            tma_load_3d(dst, &tmap_res, t.cot * CW, t.t0 + nb, t.b, &in_full[slot]);
            if (NIN == 2) tma_load_3d(dst + CB * CW * 4, &tmap_acc, t.cot * CW, t.t0 + nb, t.b, &in_full[slot]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue math warps 2..9
    const int g = warp % 4;                       // TMEM lane group of this warp
    const int half = (warp - 2) >> 2;             // the two warps of a lane group split the columns
    const int lane0 = (g * 32) % p.LR;            // first channel (within the tile) held by this warp's lanes
    const int replica = (g * 32) / p.LR;
    const int ch = lane0 + lane;
    const bool lane_ok = ch < CW;
    const bool warp_ok = lane0 < CW;
    const int colbase = (replica * 2 + half) * NCOL;   // this warp's columns inside a block
    const float sc = p.scale;
    const uint32_t off0 = (uint32_t)(colbase * CW + ch);
    const uint32_t istep = (uint32_t)CW * 4u, ostep = (uint32_t)CW * (BF16OUT ? 2u : 4u);
    const uint32_t acc_off = (uint32_t)(CB * CW) * 4u;
    const uint32_t in_base = smem_u32(in_st) + off0 * 4u;
    const uint32_t out_base = smem_u32(out_st) + off0 * (BF16OUT ? 2u : 4u);

    uint32_t acc = 0, accph = 0, cb = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const U2Tile t = u2_tile(p, tile);
      const float bv = (p.bias && lane_ok) ? BVG_LDG(p.bias + (int64_t)t.b * p.bias_bs + t.cot * CW + ch) : 0.f;

      mbar_wait(&t_full[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + acc * 256u + (uint32_t)colbase;
      for (int nb = 0; nb < t.nb_end; nb += CB, ++cb) {
        const uint32_t slot = cb % U2_IN_SLOTS, ob = cb & 1u;
        uint32_t v[NCOL];
        if (warp_ok) {
          if (NCOL == 16) tmem_ld_32x16(taddr + nb, reinterpret_cast<uint32_t(&)[16]>(v));
          else tmem_ld_32x8(taddr + nb, reinterpret_cast<uint32_t(&)[8]>(v));
        }
        float rv[NCOL], av[NCOL];
        if (NIN >= 1) {
          mbar_wait(&in_full[slot], (cb / U2_IN_SLOTS) & 1u);
          if (lane_ok) {
            const uint32_t ia = in_base + slot * U2_SLOT_BYTES;
#pragma unroll
            for (int i = 0; i < NCOL; ++i) rv[i] = ld_shared_f32(ia + i * istep);
            if (NIN == 2) {
#pragma unroll
              for (int i = 0; i < NCOL; ++i) av[i] = ld_shared_f32(ia + acc_off + i * istep);
            }
            // the loads above are only ISSUED at this point: their values must be in registers before the slot goes back to
            // the TMA producer (common.cuh, loads_landed)
            uint32_t h = fold_bits(rv);
            if (NIN == 2) h ^= fold_bits(av);
            loads_landed(sink_u32, h);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&in_free[slot]);   // values are in registers: the slot may be refilled
        }
        if (warp_ok) {
          // same for the accumulator columns before the accumulator can be handed back to the MMA warp below
          tmem_ld_wait();
          loads_landed(sink_u32, fold_bits(v));
        }
        if (nb + CB >= t.nb_end) {           // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
        mbar_wait(&out_free[ob], ((cb >> 1) & 1u) ^ 1u);
        if (lane_ok) {
          const uint32_t oa = out_base + ob * U2_SLOT_BYTES;
#pragma unroll
          for (int i = 0; i < NCOL; ++i) {
            float y = __uint_as_float(v[i]) + bv;
            if (NIN >= 1) y += rv[i];
            y *= sc;
            if (NIN == 2) y += av[i];
            if (BF16OUT) st_shared_b16(oa + i * ostep, __bfloat16_as_ushort(__float2bfloat16_rn(y)));
            else st_shared_f32(oa + i * ostep, y);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_ready[ob]);
      }
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host side ----
static int u2_nin(const ConvArgs& a) { return a.res ? (a.accum ? 2 : 1) : (a.accum ? -1 : 0); }

static bool u2_plan(const ConvArgs& a, U2Params& p) {
  const int nin = u2_nin(a);
  if (nin < 0) return false;
  if (a.in_dtype != BVG_BF16 || a.w_dtype != BVG_BF16) return false;
  if (a.Cin_p % 8 != 0 || a.Cout_r % 128 != 0 || a.Cout_n <= 0) return false;
  if (a.T <= 0 || a.T > 0x3fffffffLL || a.B <= 0) return false;
  const int halo = (a.k - 1) * a.dil;
  if (halo > 64) return false;
  const int es = a.out_dtype == BVG_BF16 ? 2 : 4;
  const int ncot = (int)ceil_div(a.Cout_n, 128);
  if (a.Cout_n % ncot) return false;
  const int CW = a.Cout_n / ncot;
  if ((CW * es) % 16 || (CW * 4) % 16) return false;
  if (((int64_t)a.out_ld * es) % 16 || ((int64_t)a.out_ld * 4) % 16 || ((int64_t)a.Cin_p * 2) % 16) return false;
  uintptr_t al = reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.out);
  if (a.res) al |= reinterpret_cast<uintptr_t>(a.res);
  if (a.accum) al |= reinterpret_cast<uintptr_t>(a.accum);
  if (al & 15) return false;

  p.bias = a.bias; p.bias_bs = a.bias_bs; p.scale = a.scale;
  p.B = a.B; p.T = (int)a.T;
  p.Cin_p = a.Cin_p; p.nchunks = (int)ceil_div(a.Cin_p, 64);
  p.k = a.k; p.dil = a.dil; p.center = (a.k - 1) / 2;
  p.CW = CW; p.n_cotiles = ncot;
  p.rep = CW <= 32 ? 4 : (CW <= 64 ? 2 : 1);
  p.LR = 128 / p.rep;
  // replicated layers: the 128-row tile already holds the copies (pack_conv_weight)
  if (p.rep > 1 && weight_replica_rows(a.Cout_n, a.Cout_r) != p.LR) return false;
  p.wrows = p.rep == 1 ? round_up(CW, 8) : 128;
  if ((int64_t)(ncot - 1) * CW + p.wrows > a.Cout_r) return false;
  p.a_stages = nin ? U2_AIN_SLOTS - U2_IN_SLOTS : U2_AIN_SLOTS;
  p.w_resident = (ncot == 1 && a.k * p.nchunks <= p.a_stages) ? 1 : 0;
  const int ncol = nin == 2 ? 8 : 16;
  p.CB = ncol * 2 * p.rep;
  if (p.CB * CW * 4 * (nin ? nin : 1) > U2_SLOT_BYTES) return false;

  // tile width: minimise (rounds over the SMs) x (cycles per tile), see DESIGN.md section 3
  const int sms = umma_sm_count();
  const int nt_max = ((U2_MAX_X_ROWS - 16 - halo) < 256 ? (U2_MAX_X_ROWS - 16 - halo) : 256) / p.CB * p.CB;
  if (nt_max < p.CB) return false;
  const double ksteps = (double)a.k * (a.Cin_p / 16);
  double best = 1e30;
  int best_nt = p.CB;
  for (int nt = p.CB; nt <= nt_max; nt += p.CB) {
    // measured (u2dbg, 384/768-channel layers): an MMA with N in (128, 256] takes ~150 cycles in the running
    // pipeline whatever N is (operand fetch + TMA fill share the shared-memory ports), so wide tiles win
    const double t_mma = ksteps * (nt > 128 ? 150.0 : (nt / 2.0 > 90.0 ? nt / 2.0 : 90.0));
    const double t_mem = nt * ((double)CW * (es + 4 * nin) + (double)a.Cin_p * 2 / ncot) / 22.0;
    const double t_epi = (double)(nt / p.CB) * (ncol * 8 + 250);
    double t = t_mma > t_mem ? t_mma : t_mem;
    t = (t > t_epi ? t : t_epi) + 500.0;
    const double rounds = (double)ceil_div((int64_t)a.B * ceil_div(a.T, nt) * ncot, sms);
    const double cost = rounds * t;
    if (cost <= best) { best = cost; best_nt = nt; }
  }
  p.NT = best_nt;
  p.x_nbox = (p.NT + halo) <= 256 ? 1 : 2;
  p.x_box_rows = round_up((p.NT + halo + p.x_nbox - 1) / p.x_nbox, 8);
  if (p.x_nbox * p.x_box_rows > U2_MAX_X_ROWS) return false;
  p.n_ttiles = (int)ceil_div(a.T, p.NT);
  p.n_tiles = (int64_t)a.B * p.n_ttiles * ncot;
  return true;
}

bool conv_umma2_supported(const ConvArgs& a) {
  U2Params p;
  return u2_plan(a, p);
}

int conv_umma2_launch(const ConvArgs& a, int variant, cudaStream_t st) {
  (void)variant;
  if (a.B <= 0 || a.T <= 0) return BVG_OK;
  U2Params p;
  if (!u2_plan(a, p)) BVG_FAIL(BVG_EINVAL, "conv_umma2: unsupported layer shape/dtype");
  const int nin = u2_nin(a);
  const int es = a.out_dtype == BVG_BF16 ? 2 : 4;
  CUtensorMap mx, mw, mo, mr, ma;
  int rc = make_map_any(&mx, a.in, 2, (uint64_t)a.Cin_p, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.Cin_p, 64,
                        (uint32_t)p.x_box_rows, 1, 128);
  if (rc) return rc;
  rc = make_map_any(&mw, a.w, 2, (uint64_t)a.Cin_p, (uint64_t)a.Cout_r, (uint64_t)a.k, (uint64_t)a.Cin_p, 64,
                    (uint32_t)p.wrows, 1, 128);
  if (rc) return rc;
  rc = make_map_any(&mo, a.out, es, (uint64_t)a.Cout_n, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.out_ld,
                    (uint32_t)p.CW, (uint32_t)p.CB, 1, 0);
  if (rc) return rc;
  mr = mo; ma = mo;
  if (nin >= 1) {
    rc = make_map_any(&mr, a.res, 4, (uint64_t)a.Cout_n, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.out_ld,
                      (uint32_t)p.CW, (uint32_t)p.CB, 1, 0);
    if (rc) return rc;
  }
  if (nin == 2) {
    rc = make_map_any(&ma, a.accum, 4, (uint64_t)a.Cout_n, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.out_ld,
                      (uint32_t)p.CW, (uint32_t)p.CB, 1, 0);
    if (rc) return rc;
  }
  const int sms = umma_sm_count();
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  const int smem = a.own_sm ? U2_SMEM_BYTES : U2_SMEM_USED;
#define BVG_U2_LAUNCH(N, O)                                                                                          \
  do {                                                                                                               \
    static std::atomic<unsigned long long> attr_done_{0};                                                              \
    if (int rc_ = smem_attr_once(conv_umma2_kernel<N, O>, U2_SMEM_BYTES, attr_done_)) return rc_;                      \
    launch_pdl(conv_umma2_kernel<N, O>, grid, U2_THREADS, smem, st, mx, mw, mo, mr, ma, p);                            \
  } while (0)
  const bool ob = a.out_dtype == BVG_BF16;
  if (nin == 0) { if (ob) BVG_U2_LAUNCH(0, true); else BVG_U2_LAUNCH(0, false); }
  else if (nin == 1) { if (ob) BVG_U2_LAUNCH(1, true); else BVG_U2_LAUNCH(1, false); }
  else { if (ob) BVG_U2_LAUNCH(2, true); else BVG_U2_LAUNCH(2, false); }
#undef BVG_U2_LAUNCH
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
