// Dilated Conv1d / polyphase ConvTranspose1d as an implicit GEMM on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   D[co, t] = sum_tap sum_ci  Wp[tap][co][ci] * X[t + (tap-center)*dil][ci]
//
//   A operand (M = 128 out-channels, K-major)  : weight tile Wp[tap][co0:co0+128][ci0:ci0+KC]   (TMA, per tap)
//   B operand (N = NT time rows,    K-major)  : activation tile X[t0-halo : t0+NT+halo][ci0:ci0+KC], channels-last
//   D (fp32, TMEM)                             : 128 lanes (co) x NT columns (t), double buffered (2*NT <= 512 cols)
//
// The activation tile (the big operand) is loaded ONCE per input-channel chunk and
// reused by all k taps: tap j reads the same shared-memory tile through a UMMA
// descriptor whose start address is advanced by j*dil rows.  Rows are
// (KC*2)-byte swizzle rows, the hardware swizzle is a function of the absolute
// shared-memory address, so a whole-row shift keeps TMA's write pattern and
// UMMA's read pattern consistent.  TMA zero-fills rows outside [0,T) of the
// 3-D tensor (C, T, B), which is exactly the conv's zero padding, per utterance.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles):
//   warp 0    : TMA producer (one lane)         -> x_full/x_empty, a_full/a_empty mbarrier rings
//   warp 1    : TMEM allocator + MMA issuer (one lane), tcgen05.commit frees stages / publishes D
//   warps 2-5 : epilogue: tcgen05.ld -> (+bias, +residual, *scale, +accum) -> coalesced global stores
//
// reference semantics: torch Conv1d/ConvTranspose1d as built in bigvgan.py:59-66,76-83,285-287,306-312.
#include "umma_common.cuh"

namespace bvg {

constexpr int UM_THREADS = 192;
constexpr int UM_M = 128;
constexpr int UM_A_STAGES = 6;
constexpr int UM_X_STAGES = 2;
constexpr int UM_A_STAGE_BYTES = UM_M * 128;      // 128 rows x (<=128 B)
constexpr int UM_MAX_X_ROWS = 320;                // NT + (k-1)*dil rounded to 2 boxes of <=160 rows
constexpr int UM_X_STAGE_BYTES = UM_MAX_X_ROWS * 128;
constexpr int UM_EPI_BUF_BYTES = 32 * UM_M * 4;     // one staged [32 time rows][<=128 channels] fp32 block
constexpr int UM_SMEM_BYTES = 1024 /*align slack*/ + UM_A_STAGES * UM_A_STAGE_BYTES + UM_X_STAGES * UM_X_STAGE_BYTES +
                              2 * UM_EPI_BUF_BYTES + 256;

struct UmmaParams {
  const float* bias;
  void* out;
  const float* res;
  const float* accum;
  float scale;
  int out_bf16;
  int B;
  int T;
  int Cin_p, Cout_n, out_ld;
  int k, dil, center;
  int KC;            // input channels per chunk: 64 / 32 / 16
  int nchunks;       // Cin_p / KC
  int NT;            // time columns per tile (<= 256, multiple of 16)
  int x_box_rows;    // rows per TMA box of the activation tile (2 boxes per chunk)
  int n_ttiles, n_cotiles;
  int64_t n_tiles;   // B * n_ttiles * n_cotiles
  int base_mode;     // debug: 1 = put (addr>>7)&7 into the descriptor base_offset field
};

template <bool PER_TAP_X>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const UmmaParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned base (SWIZZLE_128B atoms)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_st = smem;
  unsigned char* x_st = smem + UM_A_STAGES * UM_A_STAGE_BYTES;
  float* epi_st = reinterpret_cast<float*>(x_st + UM_X_STAGES * UM_X_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_st + UM_X_STAGES * UM_X_STAGE_BYTES + 2 * UM_EPI_BUF_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + UM_A_STAGES;
  uint64_t* x_full = a_empty + UM_A_STAGES;
  uint64_t* x_empty = x_full + UM_X_STAGES;
  uint64_t* t_full = x_empty + UM_X_STAGES;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int row_bytes = p.KC * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < UM_A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < UM_X_STAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
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

  const int ntaps = p.k;
  const uint32_t a_bytes = (uint32_t)UM_M * row_bytes;
  const uint32_t xbox_bytes = (uint32_t)p.x_box_rows * row_bytes;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t as = 0, aph = 0, xs = 0, xph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int cot = (int)(tile % p.n_cotiles);
        const int64_t r = tile / p.n_cotiles;
        const int tt = (int)(r % p.n_ttiles);
        const int b = (int)(r / p.n_ttiles);
        const int t0 = tt * p.NT;
        for (int c = 0; c < p.nchunks; ++c) {
          if (!PER_TAP_X) {
            mbar_wait(&x_empty[xs], xph ^ 1);
            mbar_expect_tx(&x_full[xs], 2 * xbox_bytes);
            unsigned char* dstx = x_st + xs * UM_X_STAGE_BYTES;
            const int trow = t0 - p.center * p.dil;
            tma_load_3d(dstx, &tmap_x, c * p.KC, trow, b, &x_full[xs]);
            tma_load_3d(dstx + xbox_bytes, &tmap_x, c * p.KC, trow + p.x_box_rows, b, &x_full[xs]);
            if (++xs == UM_X_STAGES) { xs = 0; xph ^= 1; }
          }
          for (int j = 0; j < ntaps; ++j) {
            if (PER_TAP_X) {
              mbar_wait(&x_empty[xs], xph ^ 1);
              mbar_expect_tx(&x_full[xs], 2 * xbox_bytes);
              unsigned char* dstx = x_st + xs * UM_X_STAGE_BYTES;
              const int trow = t0 + (j - p.center) * p.dil;
              tma_load_3d(dstx, &tmap_x, c * p.KC, trow, b, &x_full[xs]);
              tma_load_3d(dstx + xbox_bytes, &tmap_x, c * p.KC, trow + p.x_box_rows, b, &x_full[xs]);
              if (++xs == UM_X_STAGES) { xs = 0; xph ^= 1; }
            }
            mbar_wait(&a_empty[as], aph ^ 1);
            mbar_expect_tx(&a_full[as], a_bytes);
            tma_load_3d(a_st + as * UM_A_STAGE_BYTES, &tmap_w, c * p.KC, cot * UM_M, j, &a_full[as]);
            if (++as == UM_A_STAGES) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: the whole warp walks the loop with
    // warp-uniform values (descriptors stay in uniform registers); one elected lane issues.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) |
                           ((uint32_t)(UM_M >> 4) << 24);
    // descriptor template once; per MMA only the start-address field (addr >> 4) advances
    const uint64_t desc0 = make_smem_desc(0, row_bytes, 0);
    const uint32_t tap_step = (uint32_t)(p.dil * row_bytes) >> 4;
    const int nkk = p.KC / 16;
    const uint32_t a_base = smem_u32(a_st), x_base = smem_u32(x_st);
    uint32_t as = 0, aph = 0, xs = 0, xph = 0, acc = 0, accph = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&t_empty[acc], accph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.NT;
      uint32_t first = 1;
      for (int c = 0; c < p.nchunks; ++c) {
        if (!PER_TAP_X) {
          mbar_wait(&x_full[xs], xph);
          tc_fence_after();
        }
        for (int j = 0; j < ntaps; ++j) {
          if (PER_TAP_X) {
            mbar_wait(&x_full[xs], xph);
            tc_fence_after();
          }
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint64_t da0 = desc0 + ((a_base + as * UM_A_STAGE_BYTES) >> 4);
          uint64_t db0 = desc0 + ((x_base + xs * UM_X_STAGE_BYTES) >> 4);
          if (!PER_TAP_X) db0 += (uint32_t)j * tap_step;
          if (elect_one()) {
            umma_f16_ss(d_tmem, da0, db0, idesc, first ? 0u : 1u);
            for (int kk = 1; kk < nkk; ++kk) umma_f16_ss(d_tmem, da0 + 2 * kk, db0 + 2 * kk, idesc, 1u);
            umma_commit(&a_empty[as]);
            if (PER_TAP_X) umma_commit(&x_empty[xs]);
          }
          __syncwarp();
          first = 0;
          if (++as == UM_A_STAGES) { as = 0; aph ^= 1; }
          if (PER_TAP_X) {
            if (++xs == UM_X_STAGES) { xs = 0; xph ^= 1; }
          }
        }
        if (!PER_TAP_X) {
          if (elect_one()) umma_commit(&x_empty[xs]);
          __syncwarp();
          if (++xs == UM_X_STAGES) { xs = 0; xph ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&t_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  } else {
    // -------------------------------------------------- epilogue warps 2..5
    // TMEM (lane = out-channel, column = time) -> registers -> shared [time][channel] block ->
    // 16-byte vector global accesses that are contiguous along channels (and across rows when
    // the layer has <= 128 channels).  Residual / accumulate operands are loaded before any
    // store of the same block, so the loads are not serialised behind possibly aliasing stores.
    const int g = warp % 4;               // TMEM lane group this warp may access
    const int etid = (warp - 2) * 32 + lane;   // 0..127 among the epilogue threads
    uint32_t acc = 0, accph = 0;
    uint32_t blk = 0;                     // running 32-column block counter -> staging buffer parity
    // fire-and-forget L2 prefetch of one tile's residual / accumulate rows (issued one tile ahead)
    auto prefetch_tile = [&](int64_t tile) {
      if (!p.res && !p.accum) return;
      const int cot = (int)(tile % p.n_cotiles);
      const int64_t r = tile / p.n_cotiles;
      const int tt = (int)(r % p.n_ttiles);
      const int b = (int)(r / p.n_ttiles);
      const int t0 = tt * p.NT;
      const int co0 = cot * UM_M;
      const int cv = (p.Cout_n - co0) < UM_M ? (p.Cout_n - co0) : UM_M;
      int rows = p.T - t0;
      if (rows > p.NT) rows = p.NT;
      const int lpr = (cv * 4 + 127) / 128;                  // 128-byte lines per row segment
      const int64_t base = ((int64_t)b * p.T + t0) * p.out_ld + co0;
      for (int i = etid; i < rows * lpr; i += 128) {
        const int row = i / lpr, ln = i - row * lpr;
        const int64_t off = base + (int64_t)row * p.out_ld + ln * 32;
        if (p.res) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + off));
        if (p.accum) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.accum + off));
      }
    };
    if ((int64_t)blockIdx.x < p.n_tiles) prefetch_tile(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      if (tile + gridDim.x < p.n_tiles) prefetch_tile(tile + gridDim.x);
      const int cot = (int)(tile % p.n_cotiles);
      const int64_t r = tile / p.n_cotiles;
      const int tt = (int)(r % p.n_ttiles);
      const int b = (int)(r / p.n_ttiles);
      const int t0 = tt * p.NT;
      const int co0 = cot * UM_M;
      const int cv = (p.Cout_n - co0) < UM_M ? (p.Cout_n - co0) : UM_M;   // valid channels of this tile (multiple of 16)
      const int vpr = cv >> 2;                                             // float4 vectors per time row
      const uint32_t vpr_magic = (65536u + (uint32_t)vpr - 1) / (uint32_t)vpr;  // e / vpr for e < 4096
      const bool warp_has_rows = g * 32 < cv;
      const int64_t tilebase = ((int64_t)b * p.T + t0) * p.out_ld + co0;

      mbar_wait(&t_full[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + acc * (uint32_t)p.NT;
      int nb_end = p.T - t0;                                               // valid time rows in this tile
      if (nb_end > p.NT) nb_end = p.NT;
      // columns staged per barrier round: the staging buffer holds 4096 floats, so layers with
      // few valid channels stage up to 128 time columns at once (fewer barrier + residual-latency
      // round trips per tile: 2 instead of 8 for the 32-channel stage)
      int cb = (128 / cv) * 32;
      cb = cb < 32 ? 32 : (cb > 128 ? 128 : cb);
      for (int nb = 0; nb < nb_end; nb += cb, ++blk) {
        float* sbuf = epi_st + (blk & 1) * (UM_EPI_BUF_BYTES / 4);
        int cols_here = nb_end - nb;
        if (cols_here > cb) cols_here = cb;
        if (warp_has_rows) {
          const int col = g * 32 + lane;
          for (int sub = 0; sub < cols_here; sub += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + nb + sub, v);
            tmem_ld_wait();
            if (col < cv) {   // cv is a multiple of 16: the last warp with rows may be half valid
#pragma unroll
              for (int i = 0; i < 32; ++i) sbuf[(sub + i) * cv + col] = __uint_as_float(v[i]);
            }
          }
        }
        if (nb + cb >= nb_end) {           // all TMEM reads of this tile are done: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int nvec = cb * vpr;
        const int nvalid = cols_here * vpr;
        const int64_t blkbase = tilebase + (int64_t)nb * p.out_ld;
        for (int e0 = 0; e0 < nvec && e0 < nvalid; e0 += 4 * 128) {
          float4 rv[4], av[4];
          int64_t off[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * 128 + etid;
            ok[u] = e < nvalid;
            const int row = (int)(((uint32_t)e * vpr_magic) >> 16);
            const int c4 = e - row * vpr;
            off[u] = blkbase + (int64_t)row * p.out_ld + c4 * 4;
            rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            av[u] = rv[u];
            if (ok[u] && p.res) rv[u] = *reinterpret_cast<const float4*>(p.res + off[u]);
            if (ok[u] && p.accum) av[u] = *reinterpret_cast<const float4*>(p.accum + off[u]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            const int e = e0 + u * 128 + etid;
            const int row = (int)(((uint32_t)e * vpr_magic) >> 16);
            const int c4 = e - row * vpr;
            const float4 d = *reinterpret_cast<const float4*>(sbuf + e * 4);
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c4 * 4));
            float4 y;
            y.x = (d.x + bv.x + rv[u].x) * p.scale + av[u].x;
            y.y = (d.y + bv.y + rv[u].y) * p.scale + av[u].y;
            y.z = (d.z + bv.z + rv[u].z) * p.scale + av[u].z;
            y.w = (d.w + bv.w + rv[u].w) * p.scale + av[u].w;
            if (p.out_bf16) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off[u]) = pk;
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off[u]) = y;
            }
          }
        }
      }
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 3-D bf16 tensor map {d0 (fastest), d1, d2}, box {b0, b1, 1}, swizzle chosen by the box row bytes
int make_map_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                int row_bytes) {
  return make_map_4d_w(m, base, d0, d1, d2, b0, b1, 1, row_bytes);
}

// same with a box that is `b2` deep in the slowest dimension (several taps per TMA copy)
int make_map_4d_w(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                  uint32_t b2, int row_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) BVG_FAIL(BVG_ENODEV, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                  : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    BVG_FAIL(BVG_ECUDA, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu) box=(%u,%u)", (int)r,
             (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1);
  return BVG_OK;
}

// general 3-D map: element size `es` (2 = bf16, 4 = fp32), swizzle span in bytes (0 = none, 32/64/128)
int make_map_any(CUtensorMap* m, const void* base, int es, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) BVG_FAIL(BVG_ENODEV, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld_elems * (uint64_t)es, ld_elems * d1 * (uint64_t)es};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    BVG_FAIL(BVG_ECUDA, "cuTensorMapEncodeTiled failed (%d) es=%d dims=(%llu,%llu,%llu) ld=%llu box=(%u,%u,%u) sw=%d",
             (int)r, es, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)ld_elems, b0, b1, b2, swizzle_bytes);
  return BVG_OK;
}

int umma_sm_count() {
  // per device (a process may drive several); cached after the first query
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static bool conv_umma1_supported(const ConvArgs& a) {
  if (a.in_dtype != BVG_BF16 || a.w_dtype != BVG_BF16) return false;
  if (a.bias_bs != 0) return false;   // per-utterance bias: second-generation kernel only
  if (a.Cin_p % 16 != 0 || a.Cout_r % UM_M != 0) return false;
  if (a.T <= 0 || a.T > 0x7fffffffLL / 2) return false;
  if ((a.k - 1) * a.dil > 64) return false;  // activation-tile halo budget (UM_MAX_X_ROWS)
  if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.w)) & 15) return false;
  return true;
}

bool conv_umma_supported(const ConvArgs& a) { return conv_umma2_supported(a) || conv_umma1_supported(a); }

int conv_umma_launch(const ConvArgs& a, int variant, cudaStream_t st) {
  if (a.B <= 0 || a.T <= 0) return BVG_OK;
  if (!(variant & 8) && conv_umma2_supported(a)) return conv_umma2_launch(a, variant, st);   // v2 kernel (default)
  if (!conv_umma1_supported(a)) BVG_FAIL(BVG_EINVAL, "conv_umma: unsupported layer shape/dtype");
  if (a.Cout_n <= 128 && !(variant & 4) && conv_umma_t_fits(a)) return conv_umma_t_launch(a, variant, st);   // time-major variant
  UmmaParams p;
  p.bias = a.bias; p.out = a.out; p.res = a.res; p.accum = a.accum; p.scale = a.scale;
  p.out_bf16 = a.out_dtype == BVG_BF16;
  p.B = a.B; p.T = (int)a.T; p.Cin_p = a.Cin_p; p.Cout_n = a.Cout_n; p.out_ld = a.out_ld;
  p.k = a.k; p.dil = a.dil; p.center = (a.k - 1) / 2;
  p.KC = (a.Cin_p % 64 == 0) ? 64 : (a.Cin_p % 32 == 0 ? 32 : 16);
  p.nchunks = a.Cin_p / p.KC;
  p.NT = 256;
  const bool per_tap = (variant & 2) != 0;
  const int xrows = per_tap ? p.NT : p.NT + (a.k - 1) * a.dil;
  p.x_box_rows = round_up((xrows + 1) / 2, 8);
  if (2 * p.x_box_rows > UM_MAX_X_ROWS) BVG_FAIL(BVG_EINVAL, "conv_umma: halo too large");
  p.n_ttiles = (int)ceil_div(a.T, p.NT);
  p.n_cotiles = (int)ceil_div(a.Cout_n, UM_M);
  p.n_tiles = (int64_t)a.B * p.n_ttiles * p.n_cotiles;
  p.base_mode = variant & 1;

  CUtensorMap mx, mw;
  const int row_bytes = p.KC * 2;
  int rc = make_map_3d(&mx, a.in, (uint64_t)a.Cin_p, (uint64_t)a.T, (uint64_t)a.B, (uint32_t)p.KC,
                       (uint32_t)p.x_box_rows, row_bytes);
  if (rc) return rc;
  rc = make_map_3d(&mw, a.w, (uint64_t)a.Cin_p, (uint64_t)a.Cout_r, (uint64_t)a.k, (uint32_t)p.KC, UM_M, row_bytes);
  if (rc) return rc;

  const int g_sm_count = umma_sm_count();
  BVG_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_BYTES));
  BVG_CUDA(cudaFuncSetAttribute(conv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_BYTES));
  const unsigned grid = (unsigned)(p.n_tiles < g_sm_count ? p.n_tiles : g_sm_count);
  if (per_tap)
    conv_umma_kernel<true><<<grid, UM_THREADS, UM_SMEM_BYTES, st>>>(mx, mw, p);
  else
    conv_umma_kernel<false><<<grid, UM_THREADS, UM_SMEM_BYTES, st>>>(mx, mw, p);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
