// fp32-accumulate SIMT dilated Conv1d on channels-last data.  This is the
// "fp32 kernel mode" of the vocoder (<=1e-5 relative to the fp32 reference) and
// the on-GPU cross-check of the tcgen05 kernel; it is not the throughput path.
//
//   out[b,t,co] = scale*(bias[co] + sum_j sum_ci Wp[j][co][ci] * in[b, t+(j-(k-1)/2)*dil, ci] + res[b,t,co]) + accum[b,t,co]
// zero outside [0,T)  (torch Conv1d with padding (k-1)*dil/2; bigvgan.py:59-66,76-83)
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
  if (a.in_dtype == BVG_F32 && a.w_dtype == BVG_F32) return launch_simt<float, float>(a, st);
  if (a.in_dtype == BVG_BF16 && a.w_dtype == BVG_BF16) return launch_simt<bf, bf>(a, st);
  if (a.in_dtype == BVG_BF16 && a.w_dtype == BVG_F32) return launch_simt<bf, float>(a, st);
  if (a.in_dtype == BVG_F32 && a.w_dtype == BVG_BF16) return launch_simt<float, bf>(a, st);
  BVG_FAIL(BVG_EDTYPE, "conv_simt: unsupported dtypes");
}

}  // namespace bvg
