// Micro-probe: FP32 FMA issue rate on this GPU: scalar FFMA (3 distinct regs), FFMA with a
// constant-bank operand, and packed fma.rn.f32x2 (FFMA2).  Prints lane-FMA/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_ffma(float* out, float a, float b) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_3reg(float* out, float a, float b) {
  float x[16], y[16];
  for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x + i; y[i] = a + i; }
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(y[i], y[(i + 5) & 15], x[i]);
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[16], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  for (int i = 0; i < 16; ++i) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(f)); }
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  float s = 0;
  for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float* out, float a) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __cosf(x[i]);
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 4; ++which) {
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_ffma<<<sms * 8, 256>>>(out, 1.0001f, 0.5f);
      if (which == 1) k_ffma_3reg<<<sms * 8, 256>>>(out, 1.0001f, 0.5f);
      if (which == 2) k_ffma2<<<sms * 8, 256>>>(out, 1.0001f, 0.5f);
      if (which == 3) k_mufu<<<sms * 8, 256>>>(out, 1.0f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)sms * 8 * 256 * ITER * (which == 3 ? 8 : 16) * (which == 2 ? 2 : 1);
    const char* names[] = {"FFMA (reg,const,const)", "FFMA (3 distinct regs)", "FFMA2 f32x2", "cos.approx (FMUL+MUFU)"};
    printf("%-26s %.3f ms  %.1f G lane-op/s  = %.1f lane-op/clk/SM at %d MHz\n", names[which], best, ops / best / 1e6,
           ops / (best * 1e-3) / sms / (clk * 1e3), clk / 1000);
  }
  return 0;
}
