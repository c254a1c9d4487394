// Fused anti-aliased activation on the reference layout [B, C, T] (time fastest).
// Replaces the reference's `anti_alias_activation_cuda.forward`
// (alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-246) with the exact
// edge semantics of the torch operator (alias_free_activation/torch/act.py:25-30).
//
// One CTA = one time tile of one (b,c) row:
//   1. the tile plus an 8-sample halo is staged in shared memory by ONE bulk
//      async copy (cp.async.bulk / UBLKCP, mbarrier completion) when the row
//      pitch is 16-byte aligned, else by cooperative scalar loads;
//   2. each thread walks an odd-length segment (odd stride => conflict-free
//      LDS/STS) with the same rotating 6+12 register window as the
//      channels-last kernel: 24 FMA + 2 snake per output, no recomputation
//      inside a segment;
//   3. outputs are staged in shared memory and leave with ONE bulk async store
//      (or cooperative stores on the unaligned path).
#include "common.cuh"

namespace bvg {

constexpr int kBctThreads = 128;
constexpr int kBctMaxSeg = 61;   // 6n-5
constexpr int kBctHalo = 8;      // >= 5, multiple of 8 elements => 16 B for bf16, 32 B for fp32

template <typename T, bool FAST>
__global__ void __launch_bounds__(kBctThreads)
act1d_bct_kernel(T* __restrict__ dst, const T* __restrict__ src, const float* __restrict__ alpha_log,
                 const float* __restrict__ beta_log, const Taps taps, int C, int64_t Tlen, int L,
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

  const float a = expf(__ldg(alpha_log + c));
  const float ib = 1.0f / (expf(__ldg(beta_log + c)) + 1e-9f);

  __syncthreads();  // makes the mbarrier init (or the cooperative loads) visible
  if (aligned) mbar_wait(&bar, 0);

  const int64_t t0 = tile_t0 + (int64_t)threadIdx.x * L;
  int64_t t1 = t0 + L;
  const int64_t tile_end = tile_t0 + tile_len < Tlen ? tile_t0 + tile_len : Tlen;
  if (t1 > tile_end) t1 = tile_end;
  const int64_t tlast = Tlen - 1;

  if (t0 < t1) {
    float X[6], V[12];
    if (t0 >= 5 && t0 + L + 4 <= tlast && t1 - t0 == L) {
      // ---- interior segment (L = 6n-5 -> L+5 steps = n bodies of 6): no clamps/edge rules/predicates ----
      const T* ip = s_in + (t0 - 5 - lo);
      T* op = s_out + (t0 - tile_t0);
#pragma unroll
      for (int i = 0; i < 5; ++i) X[i] = to_f32<T>(ip[i]);
      ip += 5;
#pragma unroll
      for (int s = 0; s < 6; ++s) {   // first body: 5 warm-up steps + 1 full step
        X[(s + 5) % 6] = to_f32<T>(ip[s]);
        float uo = 0.f, ue = 0.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float xv = X[(s + 5 - q) % 6];
          uo = fmaf(taps.up[2 * q], xv, uo);
          ue = fmaf(taps.up[2 * q + 1], xv, ue);
        }
        V[(2 * s + 10) % 12] = snake_eval<FAST>(uo, a, ib);
        V[(2 * s + 11) % 12] = snake_eval<FAST>(ue, a, ib);
        if (s == 5) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 12; ++k) acc = fmaf(taps.down[k], V[(2 * s + k) % 12], acc);
          op[0] = from_f32<T>(acc);
        }
      }
      ip += 6;
      op += 1;
      const int nbody = (L + 5) / 6 - 1;
      for (int it = 0; it < nbody; ++it) {
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          X[(s + 5) % 6] = to_f32<T>(ip[s]);
          float uo = 0.f, ue = 0.f;
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const float xv = X[(s + 5 - q) % 6];
            uo = fmaf(taps.up[2 * q], xv, uo);
            ue = fmaf(taps.up[2 * q + 1], xv, ue);
          }
          V[(2 * s + 10) % 12] = snake_eval<FAST>(uo, a, ib);
          V[(2 * s + 11) % 12] = snake_eval<FAST>(ue, a, ib);
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 12; ++k) acc = fmaf(taps.down[k], V[(2 * s + k) % 12], acc);
          op[s] = from_f32<T>(acc);
        }
        ip += 6;
        op += 6;
      }
    } else {
    // ---- generic segment: touches a row end or is short ----
    float vend = 0.f;
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
        // past the end of this thread's segment the index may leave the staged
        // range; those steps produce nothing, so any in-range sample will do
        int idx = (int)(tl - lo);
        idx = idx < n_in ? idx : n_in - 1;
        X[(s + 5) % 6] = to_f32<T>(s_in[idx]);
        float uo = 0.f, ue = 0.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float xv = X[(s + 5 - q) % 6];
          uo = fmaf(taps.up[2 * q], xv, uo);
          ue = fmaf(taps.up[2 * q + 1], xv, ue);
        }
        float vo = snake_eval<FAST>(uo, a, ib);
        float ve = snake_eval<FAST>(ue, a, ib);
        if (t >= Tlen - 3) {
          if (t == Tlen - 3) vend = vo;
          vo = vend;
          ve = vend;
        }
        V[(2 * s + 10) % 12] = vo;
        V[(2 * s + 11) % 12] = ve;
        if (s == 2 && base == 0 && t0 == 0) {
          const float v0 = V[3];
          V[10] = v0;
          V[11] = v0;
          V[0] = v0;
          V[1] = v0;
          V[2] = v0;
        }
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 12; ++k) acc = fmaf(taps.down[k], V[(2 * s + k) % 12], acc);
        if (t >= t0 && t < t1) s_out[t - tile_t0] = from_f32<T>(acc);
      }
    }
    }
  }

  const int n_out = (int)(tile_end - tile_t0);
  if (aligned) {
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0 && n_out > 0) {
      // n_out*sizeof(T) is a multiple of 16: tile_len is a multiple of 128 and Tlen*sizeof(T) % 16 == 0
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
  // tiles: as few as possible per row, then the shortest odd segment that covers them
  const int max_tile = kBctThreads * kBctMaxSeg;
  const int tiles_per_row = (int)ceil_div(Tlen, max_tile);
  const int64_t per_tile = ceil_div(Tlen, tiles_per_row);
  int L = (int)ceil_div(per_tile, kBctThreads);
  L = (int)ceil_div(L + 5, 6) * 6 - 5;  // 6n-5: whole 6-step bodies; odd => conflict-free shared-memory walk
  if (L < 7) L = 7;                 // only the first segment of a row may see v[m<0] (needs 2*L-5 >= 0)
  if (L > kBctMaxSeg) L = kBctMaxSeg;
  const int tile_len = kBctThreads * L;   // multiple of 128 elements
  const int64_t blocks = rows * tiles_per_row;
  if (blocks > 0x7fffffffLL) BVG_FAIL(BVG_EINVAL, "act1d: tensor too large (%lld blocks)", (long long)blocks);
  const int aligned = ((Tlen * (int64_t)sizeof(T)) % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const int in_bytes = ((tile_len + 2 * kBctHalo) * (int)sizeof(T) + 127) & ~127;
  const int smem = in_bytes + tile_len * (int)sizeof(T);
  auto kern = act1d_bct_kernel<T, FAST>;
  if (smem > 48 * 1024)  // per device/context attribute; cheap enough to set on every large launch
    BVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<(unsigned)blocks, kBctThreads, smem, st>>>((T*)dst, (const T*)src, alpha_log, beta_log, taps, C,
                                                    Tlen, L, tile_len, tiles_per_row, aligned);
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
