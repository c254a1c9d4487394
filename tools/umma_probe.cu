// Micro-probe: cycles per tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) as a function of N, of the
// K-major swizzle row width (32/64/128 B) and of the row shift applied to the A or B start address
// (the conv kernels address tap j of a staged activation tile by shifting the descriptor by j*dil rows).
// One CTA per SM, one thread issues `taps * nkk` MMAs per round back to back, commit, wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../voice-tts_b200/csrc/common.cuh"
using namespace bvg;

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, int row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Case { int N, row_bytes, taps, dil, shift_a, rounds, M; };

__global__ void __launch_bounds__(128, 1) probe(Case c, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
    const uint64_t da0 = mk_desc(a_addr, c.row_bytes), db0 = mk_desc(b_addr, c.row_bytes);
    const int nkk = c.row_bytes / 32;
    const uint32_t step = (uint32_t)(c.dil * c.row_bytes) >> 4;
    uint32_t ph = 0;
    long long best = 1ll << 60, tot = 0;
    for (int r = 0; r < c.rounds; ++r) {
      const long long t0 = clock64();
      for (int rep = 0; rep < 8; ++rep)
      for (int j = 0; j < c.taps; ++j) {
        const uint64_t da = da0 + (c.shift_a ? j * step : 0u);
        const uint64_t db = db0 + (c.shift_a ? 0u : j * step);
        for (int kk = 0; kk < nkk; ++kk) umma_f16_ss(tm, da + 2 * kk, db + 2 * kk, idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t1 = clock64();
      if (r > 0) { tot += t1 - t0; if (t1 - t0 < best) best = t1 - t0; }
    }
    out[2 * blockIdx.x] = best;
    out[2 * blockIdx.x + 1] = tot / (c.rounds - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* out;
  cudaMalloc(&out, sms * 2 * sizeof(long long));
  std::vector<long long> h(sms * 2);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (argc > 1 && argv[1][0] == 'm') {
    // M = 64 against M = 128 (cta_group::1, 128-byte rows, no shift): does a half-height instruction cost less?
    printf("%5s %5s | %11s %12s\n", "M", "N", "cyc/MMA avg", "MAC/clk/SM");
    for (int M : {128, 64})
      for (int N : {256, 192, 128, 64}) {
        Case c{N, 128, 11, 0, 1, 9, M};
        probe<<<sms, 128, smem>>>(c, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s (M=%d N=%d)\n", cudaGetErrorString(e), M, N); return 1; }
        cudaMemcpy(h.data(), out, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long avg = 0;
        for (int i = 0; i < sms; ++i) avg += h[2 * i + 1];
        const double cavg = (double)avg / sms / (8 * 11 * 4);
        printf("%5d %5d | %11.1f %12.0f\n", M, N, cavg, (double)M * N * 16 / cavg);
      }
    return 0;
  }
  printf("%5s %9s %5s %4s %7s | %10s %10s %12s %10s\n", "N", "row_bytes", "taps", "dil", "shifted", "cyc/MMA min", "cyc/MMA avg", "MAC/clk/SM", "smemB/clk");
  const int Ns[] = {256, 192, 128, 96, 64, 48, 32, 16};
  const int RBs[] = {128, 64, 32};
  for (int shift_a = 0; shift_a < 2; ++shift_a)
    for (int rb : RBs)
      for (int N : Ns)
        for (int dil : {0, 1, 3, 4, 5, 8}) {
          if (shift_a == 0 && N != 256 && N != 128) continue;   // B-shift (channel-major) only for wide time tiles
          Case c{N, rb, 11, dil, shift_a, 9, 128};
          probe<<<sms, 128, smem>>>(c, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s (N=%d rb=%d dil=%d)\n", cudaGetErrorString(e), N, rb, dil); return 1; }
          cudaMemcpy(h.data(), out, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
          long long best = 1ll << 60, avg = 0;
          for (int i = 0; i < sms; ++i) { if (h[2 * i] < best) best = h[2 * i]; avg += h[2 * i + 1]; }
          const int nm = 8 * 11 * (rb / 32);
          const double cmin = (double)best / nm, cavg = (double)avg / sms / nm;
          printf("%5d %9d %5d %4d %7s | %10.1f %10.1f %12.0f %10.1f\n", N, rb, 11, dil, shift_a ? "A" : "B", cmin, cavg,
                 128.0 * N * 16 / cavg, (128 + N) * 32.0 / cavg);
        }
  return 0;
}
