// Probe for the fused AMP-unit kernel: can ordinary threads WRITE a K-major SWIZZLE_128B
// bf16 operand tile into shared memory (st.shared + fence.proxy.async) that tcgen05.mma then consumes - also through the
// row-shifted descriptors the conv kernels use for their taps?  (Today both operands arrive by TMA, which applies the swizzle
// itself; the fused kernel must produce the conv operands from registers: a1 = act(x) and a2 = act(conv1 accumulator).)
//   layout written by the threads: element (row r, k) of a [rows][64] bf16 tile at a 1024-byte aligned base lives at
//     r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2          (16-byte chunk index XOR row mod 8)
//   D[m][n] = sum_k A[m][k] * B[n + shift][k],  M = 128, N = 64, K = 64 (4 K steps), integer-valued bf16 data => exact.
// Also times how fast 8 warps can write such a tile in the epilogue pattern (lane = channel, one row per step).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/swz_probe tools/swz_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../voice-tts_b200/csrc/common.cuh"
using namespace bvg;

__device__ __forceinline__ uint64_t mk_desc128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * 128) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ inline int val_a(int m, int k) { return ((m * 7 + k * 3) % 9) - 4; }
__host__ __device__ inline int val_b(int n, int k) { return ((n * 5 + k * 11) % 7) - 3; }

constexpr int M = 128, N = 64, K = 64, BROWS = 96;

__global__ void __launch_bounds__(256, 1) probe(int shift, float* out, long long* cyc) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_t = smem;                 // 128 rows x 128 B
  unsigned char* b_t = smem + 16384;         // 96 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  // A: thread-written, one element at a time (2-byte stores), epilogue pattern: lane = k (two passes of 32), rows over warps
  const long long t0 = clock64();
  for (int r = warp; r < M; r += 8)
    for (int kb = 0; kb < K; kb += 32) {
      const int k = kb + lane;
      const uint32_t off = (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2);
      st_shared_b16(smem_u32(a_t) + off, __bfloat16_as_ushort(__float2bfloat16_rn((float)val_a(r, k))));
    }
  for (int r = warp; r < BROWS; r += 8)
    for (int kb = 0; kb < K; kb += 32) {
      const int k = kb + lane;
      const uint32_t off = (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2);
      st_shared_b16(smem_u32(b_t) + off, __bfloat16_as_ushort(__float2bfloat16_rn((float)val_b(r, k))));
    }
  const long long t1 = clock64();
  fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = mk_desc128(smem_u32(a_t));
    const uint64_t db = mk_desc128(smem_u32(b_t)) + (uint64_t)((shift * 128) >> 4);   // row-shifted start address, base_offset 0
    for (int kk = 0; kk < K / 16; ++kk) umma_f16_ss(tm, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (warp < 4) {
    uint32_t v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
      tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, M * N * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  std::vector<float> h(M * N);
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int bad_total = 0;
  for (int shift : {0, 1, 2, 3, 5, 8, 15, 25}) {
    cudaMemset(out, 0, M * N * sizeof(float));
    probe<<<1, 256, smem>>>(shift, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s (shift %d)\n", cudaGetErrorString(e), shift); return 1; }
    cudaMemcpy(h.data(), out, M * N * sizeof(float), cudaMemcpyDeviceToHost);
    long long c = 0; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int k = 0; k < K; ++k) ref += val_a(m, k) * val_b(n + shift, k);
        if (h[m * N + n] != (float)ref) { if (bad < 3) printf("  mismatch m=%d n=%d got %g want %d\n", m, n, h[m * N + n], ref); ++bad; }
      }
    printf("thread-written SWIZZLE_128B operands, B shifted by %2d rows: %d / %d wrong; tile write (224 rows x 64 ch, 8 warps, 2-byte stores) %lld cycles\n",
           shift, bad, M * N, c);
    bad_total += bad;
  }
  printf(bad_total ? "SWZ PROBE FAILED\n" : "SWZ PROBE OK\n");
  return bad_total != 0;
}
