// Probe for VERDICT item 3 / SURVEY section 7 "hard parts": the anti-aliased activation with BOTH FIRs on the tensor cores
// (banded-Toeplitz matrices, mma.sync.m16n8k16 bf16 with three-term hi/lo splits for fp32-class accuracy), interior rows
// only, channels-last fp32 in -> bf16 out - timed against the shipped FFMA2 kernel (bvg_act1d_cl_fwd) on the same tensor.
//
// One warp owns 16 channels (M) and walks time in steps of 8 input rows:
//   up   : U[16 ch x 16 up-samples] = X[16 ch x 16 rows] . BU[16 x 16]     2 n-tiles x 3 split terms = 6 MMAs
//          (the 16-row window advances by 8: its upper k-half is next step's lower half - no shuffles)
//   snake: on the accumulator fragments (fp32), then hi/lo split -> the m16n8 accumulator layout IS the A-fragment layout
//   down : Y[16 ch x 8 outputs] = V[16 ch x 32 up-samples] . BD[32 x 8]    2 k-steps x 3 split terms = 6 MMAs
// => 12 MMAs per 128 elements.  Build: nvcc -arch=sm_100a -O3 -o tools/fir_mma_probe tools/fir_mma_probe.cu
//    -Lvoice-tts_b200 -lbvg_b200 -Xlinker -rpath,$PWD/voice-tts_b200       Run: tools/fir_mma_probe [C] [T] [B]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/bvg_b200.h"

struct Frag { uint32_t r[2]; };   // B fragment of m16n8k16 (bf16x2 pairs)
struct Consts {
  float fu[12];   // up taps (x2 gain included)
  float fd[12];   // down taps
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// (hi, lo) bf16x2 packs of two floats: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float up_coef(const Consts& c, int i, int n) {   // BU[i][n]: window row i (t = tb-4+i) -> up-sample 2tb+n
  const int j = n >> 1;
  const int q = (n & 1) ? j + 7 - i : j + 6 - i;
  if (q < 0 || q > 5) return 0.f;
  return (n & 1) ? c.fu[2 * q] : c.fu[2 * q + 1];
}
__device__ __forceinline__ float down_coef(const Consts& c, int i, int j) { // BD[i][j]: window sample i (m = 2t0-8+i) -> output t0+j
  const int k = i - 2 * j - 3;
  return (k < 0 || k > 11) ? 0.f : c.fd[k];
}

// x: [T][C] fp32, y: [T][C] bf16; outputs written for t in [t_lo, t_hi) (interior: t_lo >= 16, t_hi <= T - 16)
__global__ void __launch_bounds__(128, 4)
fir_mma_kernel(__nv_bfloat16* __restrict__ y, const float* __restrict__ x, const float* __restrict__ alpha_log,
               const float* __restrict__ beta_log, Consts cst, int C, int64_t T, int seg, int64_t t_lo, int64_t t_hi, int nseg) {
  const int lane = threadIdx.x % 32, g = lane / 4, q4 = lane % 4;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int cgroups = C / 16;
  if (wid >= (int64_t)cgroups * nseg) return;
  const int cg = (int)(wid % cgroups);
  const int64_t s0 = t_lo + (wid / cgroups) * seg;        // first output row of this warp's segment (multiple of 8 offset)
  int64_t s1 = s0 + seg; if (s1 > t_hi) s1 = t_hi;
  const int c0 = cg * 16 + g, c1 = c0 + 8;
  const float a0 = __expf(alpha_log[c0]), a1 = __expf(alpha_log[c1]);
  const float ib0 = 1.0f / (__expf(beta_log[c0]) + 1e-9f), ib1 = 1.0f / (__expf(beta_log[c1]) + 1e-9f);

  // constant B fragments (hi / lo): thread holds B[k = 2q4, 2q4+1 (+8)][n = g]
  uint32_t bu_h[2][2], bu_l[2][2], bd_h[2][2], bd_l[2][2];
#pragma unroll
  for (int tile = 0; tile < 2; ++tile)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k0 = 2 * q4 + 8 * h;
      split2(up_coef(cst, k0, 8 * tile + g), up_coef(cst, k0 + 1, 8 * tile + g), bu_h[tile][h], bu_l[tile][h]);
      split2(down_coef(cst, 16 * tile + k0, g), down_coef(cst, 16 * tile + k0 + 1, g), bd_h[tile][h], bd_l[tile][h]);
    }

  // step s handles input rows [tb, tb + 8) and emits outputs [tb - 4, tb + 4); the first step is a warm-up (its outputs are dropped)
  int64_t tb = s0 - 4;                       // so that the first kept chunk [tb+8-4, tb+8+4) = [s0, s0+8)
  uint32_t xa_h[4], xa_l[4];                 // A fragment of the 16-row window x[tb-4 .. tb+11]: [0],[1] = rows 2q4,2q4+1 (ch g / g+8)
  uint32_t va_h[4], va_l[4];                 // V fragment of the previous step (k-step 0 of the down MMA)
  {
    const float* p = x + (tb - 4 + 2 * q4) * C;
    split2(__ldg(p + c0), __ldg(p + C + c0), xa_h[0], xa_l[0]);
    split2(__ldg(p + c1), __ldg(p + C + c1), xa_h[1], xa_l[1]);
  }
  bool first = true;
  for (; tb - 4 < s1; tb += 8) {
    {
      const float* p = x + (tb + 4 + 2 * q4) * C;
      split2(__ldg(p + c0), __ldg(p + C + c0), xa_h[2], xa_l[2]);
      split2(__ldg(p + c1), __ldg(p + C + c1), xa_h[3], xa_l[3]);
    }
    uint32_t vb_h[4], vb_l[4];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      float u[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16(u, xa_l, bu_h[tile]);
      mma_bf16(u, xa_h, bu_l[tile]);
      mma_bf16(u, xa_h, bu_h[tile]);
      // snake: v = u + 1/(beta) * sin^2(alpha u)   (u[0],u[1]: channel c0; u[2],u[3]: channel c1)
      float s;
      s = __sinf(u[0] * a0); u[0] = fmaf(s * ib0, s, u[0]);
      s = __sinf(u[1] * a0); u[1] = fmaf(s * ib0, s, u[1]);
      s = __sinf(u[2] * a1); u[2] = fmaf(s * ib1, s, u[2]);
      s = __sinf(u[3] * a1); u[3] = fmaf(s * ib1, s, u[3]);
      split2(u[0], u[1], vb_h[2 * tile], vb_l[2 * tile]);
      split2(u[2], u[3], vb_h[2 * tile + 1], vb_l[2 * tile + 1]);
    }
    if (!first) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16(o, va_l, bd_h[0]); mma_bf16(o, va_h, bd_l[0]); mma_bf16(o, va_h, bd_h[0]);
      mma_bf16(o, vb_l, bd_h[1]); mma_bf16(o, vb_h, bd_l[1]); mma_bf16(o, vb_h, bd_h[1]);
      const int64_t t = tb - 4 + 2 * q4;
      if (t < s1) {
        y[t * C + c0] = __float2bfloat16_rn(o[0]);
        y[t * C + c1] = __float2bfloat16_rn(o[2]);
      }
      if (t + 1 < s1) {
        y[(t + 1) * C + c0] = __float2bfloat16_rn(o[1]);
        y[(t + 1) * C + c1] = __float2bfloat16_rn(o[3]);
      }
    }
    first = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) { va_h[i] = vb_h[i]; va_l[i] = vb_l[i]; }
    xa_h[0] = xa_h[2]; xa_h[1] = xa_h[3]; xa_l[0] = xa_l[2]; xa_l[1] = xa_l[3];
  }
}

// bare issue rate of the MMA used above: 8 independent accumulators per warp
__global__ void mma_rate_kernel(float* out, int iters) {
  float d[8][4] = {};
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3c003c00u, 0x3c003c00u};
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_bf16(d[i], a, b);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static void kaiser_taps(float* f) {   // kaiser_sinc_filter1d(0.25, 0.3, 12), double precision then rounded (probe only)
  const double cutoff = 0.25, half_width = 0.3;
  const double A = 2.285 * (12 / 2 - 1) * M_PI * 4 * half_width + 7.95;
  const double beta = A > 50 ? 0.1102 * (A - 8.7) : (A >= 21 ? 0.5842 * pow(A - 21, 0.4) + 0.07886 * (A - 21) : 0.0);
  auto i0 = [](double x) { double s = 1, t = 1; for (int k = 1; k < 40; ++k) { t *= (x / (2 * k)) * (x / (2 * k)); s += t; } return s; };
  double w[12], sum = 0;
  for (int n = 0; n < 12; ++n) {
    const double r = 2.0 * n / 11 - 1.0;
    const double win = i0(beta * sqrt(1 - r * r)) / i0(beta);
    const double tt = n - 6 + 0.5, xx = 2 * cutoff * tt;
    const double sinc = fabs(xx) < 1e-12 ? 1.0 : sin(M_PI * xx) / (M_PI * xx);
    w[n] = 2 * cutoff * win * sinc; sum += w[n];
  }
  for (int n = 0; n < 12; ++n) f[n] = (float)(w[n] / sum);
}

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 96;
  const int64_t T = argc > 2 ? atoll(argv[2]) : 55104;
  const int B = argc > 3 ? atoi(argv[3]) : 16;
  const int64_t TT = T * B;                       // the probe treats the batch as one long sequence (interior rows only)
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float taps[12]; kaiser_taps(taps);
  Consts cst; for (int i = 0; i < 12; ++i) { cst.fu[i] = 2.f * taps[i]; cst.fd[i] = taps[i]; }
  std::vector<float> hx((size_t)TT * C), ha(C), hb(C);
  srand(1);
  for (auto& v : hx) v = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
  for (int c = 0; c < C; ++c) { ha[c] = (rand() / (float)RAND_MAX - 0.5f); hb[c] = (rand() / (float)RAND_MAX - 0.5f); }
  float *dx, *da, *db; __nv_bfloat16 *dy, *dref;
  cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&da, C * 4); cudaMalloc(&db, C * 4);
  cudaMalloc(&dy, hx.size() * 2); cudaMalloc(&dref, hx.size() * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(da, ha.data(), C * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), C * 4, cudaMemcpyHostToDevice);
  cudaMemset(dy, 0, hx.size() * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;

  // 1. bare mma.sync rate
  { float* o; cudaMalloc(&o, (size_t)sms * 8 * 256 * 4);
    mma_rate_kernel<<<sms * 8, 256>>>(o, 256);
    cudaEventRecord(e0); mma_rate_kernel<<<sms * 8, 256>>>(o, 4096); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * 8 * 8 * 4096 * 8;
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("mma.sync m16n8k16 bf16: %.1f G MMA/s = %.1f TFLOP/s dense (%.3f MMA/clk/SM at the %d MHz max clock)\n",
           mmas / ms / 1e6, mmas * 4096 / ms / 1e9, mmas / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
    cudaFree(o); }

  // 2. shipped kernel (FFMA2 sliding windows), fp32 -> bf16, fast snake, same tensor as ONE utterance of TT rows
  float upt[12], dnt[12]; for (int i = 0; i < 12; ++i) { upt[i] = taps[i]; dnt[i] = taps[i]; }
  int rc = bvg_act1d_cl_fwd(dref, dx, da, db, upt, dnt, 1, TT, C, BVG_F32, BVG_BF16, BVG_ACT_FAST_SIN, nullptr);
  if (rc) { printf("bvg_act1d_cl_fwd failed: %s\n", bvg_last_error()); return 1; }
  cudaDeviceSynchronize();
  float best_ship = 1e9;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    bvg_act1d_cl_fwd(dref, dx, da, db, upt, dnt, 1, TT, C, BVG_F32, BVG_BF16, BVG_ACT_FAST_SIN, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best_ship) best_ship = ms;
  }
  // 3. tensor-core form
  const int64_t t_lo = 16, t_hi = (TT - 16) / 8 * 8;
  float best_mma = 1e9; int best_seg = 0;
  for (int seg : {64, 128, 256, 512}) {
    const int nseg = (int)((t_hi - t_lo + seg - 1) / seg);
    const int64_t warps = (int64_t)(C / 16) * nseg;
    const unsigned blocks = (unsigned)((warps * 32 + 127) / 128);
    fir_mma_kernel<<<blocks, 128>>>(dy, dx, da, db, cst, C, TT, seg, t_lo, t_hi, nseg);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("fir_mma_kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float b = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      fir_mma_kernel<<<blocks, 128>>>(dy, dx, da, db, cst, C, TT, seg, t_lo, t_hi, nseg);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < b) b = ms;
    }
    printf("  tensor-core form, %3d-row segments: %.3f ms\n", seg, b);
    if (b < best_mma) { best_mma = b; best_seg = seg; }
  }
  // 4. agreement on the interior
  std::vector<__nv_bfloat16> hy(hx.size()), hr(hx.size());
  cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost); cudaMemcpy(hr.data(), dref, hr.size() * 2, cudaMemcpyDeviceToHost);
  double se = 0, sr = 0, maxd = 0; size_t ndiff = 0, n = 0;
  for (int64_t t = t_lo; t < t_hi; ++t)
    for (int c = 0; c < C; ++c, ++n) {
      const double a = __bfloat162float(hy[t * C + c]), r = __bfloat162float(hr[t * C + c]);
      se += (a - r) * (a - r); sr += r * r; if (fabs(a - r) > maxd) maxd = fabs(a - r); if (a != r) ++ndiff;
    }
  const double el = (double)TT * C;
  printf("C = %d, rows = %lld (%.1f M elements), fp32 in -> bf16 out, fast snake, best of 5\n", C, (long long)TT, el / 1e6);
  printf("shipped FFMA2 kernel      : %.3f ms  %.3f T elements/s  %.0f GB/s algorithmic (6 B/element)\n", best_ship, el / best_ship / 1e9, el * 6 / best_ship / 1e6);
  printf("tensor-core FIR (seg %3d)  : %.3f ms  %.3f T elements/s  %.0f GB/s algorithmic   = %.2fx the shipped kernel\n", best_seg, best_mma,
         el / best_mma / 1e9, el * 6 / best_mma / 1e6, best_ship / best_mma);
  printf("agreement on interior rows: SNR %.1f dB, max |diff| %.4g, %.2f %% of the bf16 outputs differ (1 ulp of bf16 = 2^-8 relative)\n",
         10 * log10(sr / (se > 0 ? se : 1e-300)), maxd, 100.0 * ndiff / n);
  return 0;
}
