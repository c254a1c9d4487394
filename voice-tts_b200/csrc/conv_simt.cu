// fp32-accumulate SIMT dilated Conv1d on channels-last data.  This is the
// "fp32 kernel mode" of the vocoder (<=1e-5 relative to the fp32 reference) and
// the on-GPU cross-check of the tcgen05 kernel; it is not the throughput path.
//
//   out[b,t,co] = scale*(bias[co] + sum_j sum_ci Wp[j][co][ci] * in[b, t+(j-(k-1)/2)*dil, ci] + res[b,t,co]) + accum[b,t,co]
// zero outside [0,T)  (torch Conv1d with padding (k-1)*dil/2; bigvgan.py:59-66,76-83)
#include <cstdlib>

#include "conv.cuh"

namespace bvg {

constexpr int TS_T = 64, TS_C = 64, TS_K = 16;

template <typename Tin, typename Tw, typename Tout>
__global__ void __launch_bounds__(256)
conv_simt_kernel(ConvArgs a) {
  __shared__ __align__(16) float Xs[TS_K][TS_T + 4];
  __shared__ __align__(16) float Ws[TS_K][TS_C + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int b = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.x * TS_T;
  const int co0 = blockIdx.y * TS_C;
  const Tin* in = reinterpret_cast<const Tin*>(a.in) + (int64_t)b * a.T * a.Cin_p;
  const Tw* w = reinterpret_cast<const Tw*>(a.w);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lrow = threadIdx.x / 4;        // 0..63: time row (X) / out-channel row (W)
  const int lq = (threadIdx.x % 4) * 4;    // 0,4,8,12: first of 4 input channels
  const int center = (a.k - 1) / 2;

  for (int j = 0; j < a.k; ++j) {
    const int64_t tsrc = t0 + lrow + (int64_t)(j - center) * a.dil;
    const bool trow_ok = tsrc >= 0 && tsrc < a.T;
    const int wrow = co0 + lrow;
    const bool wrow_ok = wrow < a.Cout_r;
    for (int ci0 = 0; ci0 < a.Cin_p; ci0 += TS_K) {
      float xv[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      if (trow_ok && ci0 + lq < a.Cin_p) {
        const Tin* p = in + tsrc * a.Cin_p + ci0 + lq;
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = ld_mut_f32<Tin>(p + q);
      }
      if (wrow_ok && ci0 + lq < a.Cin_p) {
        const Tw* p = w + ((int64_t)j * a.Cout_r + wrow) * a.Cin_p + ci0 + lq;
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = to_f32<Tw>(p[q]);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        Xs[lq + q][lrow] = xv[q];
        Ws[lq + q][lrow] = wv[q];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TS_K; ++kk) {
        const float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
        const float4 wb = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        const float xr[4] = {xa.x, xa.y, xa.z, xa.w};
        const float wr[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(xr[i], wr[jj], acc[i][jj]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t t = t0 + ty * 4 + i;
    if (t >= a.T) continue;
    const int64_t rowoff = ((int64_t)b * a.T + t) * a.out_ld;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int co = co0 + tx * 4 + jj;
      if (co >= a.Cout_n) continue;
      float v = acc[i][jj];
      if (a.bias) v += BVG_LDG(a.bias + (int64_t)b * a.bias_bs + co);
      if (a.res) v += BVG_LDG(a.res + rowoff + co);
      v *= a.scale;
      if (a.accum) v += BVG_LDG(a.accum + rowoff + co);
      reinterpret_cast<Tout*>(a.out)[rowoff + co] = from_f32<Tout>(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second-generation fp32 kernel for the wide layers (the hot kernel of the <= 1e-5 parity mode): 128 time rows x TC out
// channels per block (TC = 128 or 64), 16-channel K slabs double-buffered in shared memory (the next slab's global loads are
// in flight while the current one is multiplied), 8 x TC/16 register tile per thread fed by 128-bit shared loads.  Same
// formula and the same fp32 FMA chain per output as conv_simt_kernel (taps outermost, channels ascending), so the two
// kernels agree to the last bit and the choice between them is purely a matter of tile efficiency.
#ifndef BVG_SIMT2_K
#define BVG_SIMT2_K 16
#endif
#ifndef BVG_SIMT2_MINBLK
#define BVG_SIMT2_MINBLK 2
#endif
constexpr int T2_T = 128, T2_K = BVG_SIMT2_K, T2_PAD = 4, T2_G = 20;
template <int TC>
constexpr int simt2_smem_bytes() { return (2 * T2_K * (T2_T + T2_PAD) + 2 * T2_K * (TC + T2_PAD) + (T2_T + TC) * T2_G) * 4; }

template <int TC>
__global__ void __launch_bounds__(256, BVG_SIMT2_MINBLK)
conv_simt2_kernel(ConvArgs a) {
  constexpr int NJ = TC / 16;          // out channels per thread: 8 or 4
  constexpr int NG = NJ / 4;           // float4 groups per thread
  constexpr int WLD = TC / 128 + 1;    // W slab: 16 floats of each of TC rows = TC*16/256 floats per thread / 4 = float4 loads (2 or 1)
  // dynamic shared memory: transposed operand tiles Xs[2][16][132], Ws[2][16][TC+4] (double buffered) and the cp.async
  // staging rows Xg[128][20], Wg[TC][20] in the global layout (20-float pitch: conflict-free 128-bit reads of a thread's own row)
  extern __shared__ __align__(16) float smem_f[];
  typedef float (*XsT)[T2_K][T2_T + T2_PAD];
  typedef float (*WsT)[T2_K][TC + T2_PAD];
  XsT Xs = reinterpret_cast<XsT>(smem_f);
  WsT Ws = reinterpret_cast<WsT>(smem_f + 2 * T2_K * (T2_T + T2_PAD));
  float* Xg = smem_f + 2 * T2_K * (T2_T + T2_PAD) + 2 * T2_K * (TC + T2_PAD);
  float* Wg = Xg + T2_T * T2_G;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int b = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.x * T2_T;
  const int co0 = blockIdx.y * TC;
  const float* in = reinterpret_cast<const float*>(a.in) + (int64_t)b * a.T * a.Cin_p;
  const float* w = reinterpret_cast<const float*>(a.w);

  // accumulators as f32x2 pairs over adjacent out channels: fma.rn.f32x2 is two independent round-to-nearest FMAs (the
  // result is bit-identical to scalar FMAs) but reads half as many register operands per FMA
  unsigned long long acc[8][NJ / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ / 2; ++j) acc[i][j] = 0ull;

  // loaders: X slab = 128 rows x 16 channels -> thread (row = tid % 128, 8 channels at (tid / 128) * 8): consecutive lanes own
  // consecutive rows, so the transposed shared-memory stores are conflict-free.  W slab = TC rows x 16 channels likewise.
  const int xrow = tid % 128, xc = (tid / 128) * 8;
  const int wrow = tid % TC, wc = (tid / TC) * (TC == 128 ? 8 : 4);
  const int center = (a.k - 1) / 2;
  const int nslab = a.Cin_p / T2_K;
  const int total = a.k * nslab;

  // Global -> shared with cp.async (no registers held across the FMA block, so the copies really are in flight while the
  // current slab is multiplied; as register prefetch the compiler sank the loads below the block to save registers and the
  // first shared-memory store behind it became the top stall).  Each thread stages and later transposes ITS OWN 32 (16)
  // bytes, so only cp.async.wait_group orders the staging buffer.  Rows outside [0, T) are zero-filled (src-size 0).
  const uint32_t xg_addr = (uint32_t)__cvta_generic_to_shared(Xg + xrow * T2_G + xc);
  const uint32_t wg_addr = (uint32_t)__cvta_generic_to_shared(Wg + wrow * T2_G + wc);
  auto gload = [&](int it) {
    const int j = it / nslab, ci0 = (it % nslab) * T2_K;
    const int64_t tsrc = t0 + xrow + (int64_t)(j - center) * a.dil;
    const bool ok = tsrc >= 0 && tsrc < a.T;
    const float* p = in + (ok ? tsrc : 0) * a.Cin_p + ci0 + xc;
    const int nb = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xg_addr), "l"(p), "r"(nb) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xg_addr + 16), "l"(p + 4), "r"(nb) : "memory");
    const float* q = w + ((int64_t)j * a.Cout_r + co0 + wrow) * a.Cin_p + ci0 + wc;
#pragma unroll
    for (int i = 0; i < WLD; ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wg_addr + 16 * i), "l"(q + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto sstore = [&](int buf) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const float4 x0 = *reinterpret_cast<const float4*>(Xg + xrow * T2_G + xc);
    const float4 x1 = *reinterpret_cast<const float4*>(Xg + xrow * T2_G + xc + 4);
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) Xs[buf][xc + q][xrow] = xv[q];
#pragma unroll
    for (int i = 0; i < WLD; ++i) {
      const float4 wq = *reinterpret_cast<const float4*>(Wg + wrow * T2_G + wc + 4 * i);
      Ws[buf][wc + 4 * i + 0][wrow] = wq.x;
      Ws[buf][wc + 4 * i + 1][wrow] = wq.y;
      Ws[buf][wc + 4 * i + 2][wrow] = wq.z;
      Ws[buf][wc + 4 * i + 3][wrow] = wq.w;
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) gload(it + 1);
#pragma unroll
    for (int kk = 0; kk < T2_K; ++kk) {
      const float4 xa = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * 4]);
      const float4 xb = *reinterpret_cast<const float4*>(&Xs[buf][kk][64 + ty * 4]);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      unsigned long long wv[NJ / 2];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 wq = *reinterpret_cast<const float4*>(&Ws[buf][kk][g * 64 + tx * 4]);
        asm("mov.b64 %0, {%1, %2};" : "=l"(wv[2 * g]) : "f"(wq.x), "f"(wq.y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(wv[2 * g + 1]) : "f"(wq.z), "f"(wq.w));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        unsigned long long xx;
        asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(xv[i]));
#pragma unroll
        for (int j = 0; j < NJ / 2; ++j)
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i][j]) : "l"(xx), "l"(wv[j]));
      }
    }
    if (it + 1 < total) {
      sstore(buf ^ 1);       // the other buffer was last read one iteration ago, before the barrier below
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t t = t0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (t >= a.T) continue;
    const int64_t rowoff = ((int64_t)b * a.T + t) * a.out_ld;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int co = co0 + (j / 4) * 64 + tx * 4 + (j % 4);
      if (co >= a.Cout_n) continue;
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i][j / 2]));
      float v = (j & 1) ? hi : lo;
      if (a.bias) v += BVG_LDG(a.bias + (int64_t)b * a.bias_bs + co);
      if (a.res) v += BVG_LDG(a.res + rowoff + co);
      v *= a.scale;
      if (a.accum) v += BVG_LDG(a.accum + rowoff + co);
      if (a.out_dtype == BVG_BF16) reinterpret_cast<__nv_bfloat16*>(a.out)[rowoff + co] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(a.out)[rowoff + co] = v;
    }
  }
}

static bool simt2_ok(const ConvArgs& a) {
  return a.in_dtype == BVG_F32 && a.w_dtype == BVG_F32 && a.Cin_p % T2_K == 0 && a.Cout_n > 32 && a.Cout_r % 128 == 0 &&
         (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0;
}

template <typename Tin, typename Tw>
static int launch_simt(const ConvArgs& a, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(a.T, TS_T), (unsigned)ceil_div(a.Cout_n, TS_C), (unsigned)a.B);
  if (a.out_dtype == BVG_BF16)
    conv_simt_kernel<Tin, Tw, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else
    conv_simt_kernel<Tin, Tw, float><<<grid, 256, 0, st>>>(a);
  BVG_LAUNCHED();
  return BVG_OK;
}

int conv_simt_launch(const ConvArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.T <= 0) return BVG_OK;
  if (a.Cin_p % 4 != 0) BVG_FAIL(BVG_EINVAL, "conv_simt: Cin_p=%d not a multiple of 4", a.Cin_p);
  typedef __nv_bfloat16 bf;
  static const bool old_only = getenv("BVG_SIMT_V1") != nullptr;   // A/B timing
  if (!old_only && simt2_ok(a)) {
    if (a.Cout_n > 64) {
      dim3 grid((unsigned)ceil_div(a.T, T2_T), (unsigned)ceil_div(a.Cout_n, 128), (unsigned)a.B);
      BVG_CUDA(cudaFuncSetAttribute(conv_simt2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, simt2_smem_bytes<128>()));
      conv_simt2_kernel<128><<<grid, 256, simt2_smem_bytes<128>(), st>>>(a);
    } else {
      dim3 grid((unsigned)ceil_div(a.T, T2_T), (unsigned)ceil_div(a.Cout_n, 64), (unsigned)a.B);
      BVG_CUDA(cudaFuncSetAttribute(conv_simt2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, simt2_smem_bytes<64>()));
      conv_simt2_kernel<64><<<grid, 256, simt2_smem_bytes<64>(), st>>>(a);
    }
    BVG_LAUNCHED();
    return BVG_OK;
  }
  if (a.in_dtype == BVG_F32 && a.w_dtype == BVG_F32) return launch_simt<float, float>(a, st);
  if (a.in_dtype == BVG_BF16 && a.w_dtype == BVG_BF16) return launch_simt<bf, bf>(a, st);
  if (a.in_dtype == BVG_BF16 && a.w_dtype == BVG_F32) return launch_simt<bf, float>(a, st);
  if (a.in_dtype == BVG_F32 && a.w_dtype == BVG_BF16) return launch_simt<float, bf>(a, st);
  BVG_FAIL(BVG_EDTYPE, "conv_simt: unsupported dtypes");
}

}  // namespace bvg
