// Micro-probe: is the L1 of an SM invalidated between two launches when a CTA of ANOTHER kernel stays resident on it?
//
// The multi-stream schedule of the vocoder once returned wrong samples whenever stand-alone activation blocks became
// co-resident with a persistent tcgen05 conv CTA of another stream.  Hypothesis: the per-launch L1
// invalidation does not happen on an SM that still runs a CTA of another kernel, so L1-allocating loads (ld.global.ca,
// ld.global.nc) of a buffer that an earlier launch of the reader cached on that SM, and that a kernel on other SMs has
// rewritten since, return the OLD lines; ld.global.cg (L2, the point of coherence) does not.
//
//   stream A: holder kernel, one CTA per SM with ~200 KB of shared memory, spins on clock64 for ~40 ms (no flags, no
//             waiting on other kernels); run once WITH and once WITHOUT it
//   stream B: for it = 1 .. N:  writer<<<2 blocks>>>(buf := it)   then   reader<<<8 blocks per SM>>>(count buf != it)
//             for each load kind; the 4 KB buffer is read in full by every reader block, so every SM caches every line
#include <cstdio>
#include <cuda_runtime.h>

__global__ void holder(long long cycles, int* sink) {
  extern __shared__ unsigned char sm[];
  const long long t0 = clock64();
  int acc = 0;
  while (clock64() - t0 < cycles) acc += sm[(threadIdx.x * 64) % 1024];
  if (acc == 0x7fffffff) *sink = acc;
}
__global__ void writer(int* buf, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = v;
}
template <int KIND>   // 0: ld.global.ca (default caching), 1: ld.global.nc (__ldg), 2: ld.global.cg
__global__ void reader(const int* buf, int n, int expect, unsigned long long* stale) {
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int v;
    if (KIND == 0) asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(buf + i));
    else if (KIND == 1) v = __ldg(buf + i);
    else v = __ldcg(buf + i);
    bad += v != expect;
  }
  if (bad) atomicAdd(stale, (unsigned long long)bad);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n = 1024, iters = 400;
  int *buf, *sink;
  unsigned long long* stale;
  cudaMalloc(&buf, n * sizeof(int));
  cudaMalloc(&sink, sizeof(int));
  cudaMalloc(&stale, 3 * sizeof(unsigned long long));
  cudaStream_t sa, sb;
  cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
  const int hold_smem = 200 * 1024;
  cudaFuncSetAttribute(holder, cudaFuncAttributeMaxDynamicSharedMemorySize, hold_smem);
  for (int with_holder = 0; with_holder < 2; ++with_holder) {
    cudaMemset(stale, 0, 3 * sizeof(unsigned long long));
    cudaDeviceSynchronize();
    if (with_holder) holder<<<sms, 128, hold_smem, sa>>>(120000000LL, sink);   // ~60-80 ms: outlives the loop below
    for (int it = 1; it <= iters; ++it) {
      writer<<<2, 256, 0, sb>>>(buf, n, it);
      reader<0><<<sms * 8, 128, 0, sb>>>(buf, n, it, stale + 0);
      reader<1><<<sms * 8, 128, 0, sb>>>(buf, n, it, stale + 1);
      reader<2><<<sms * 8, 128, 0, sb>>>(buf, n, it, stale + 2);
    }
    cudaStreamSynchronize(sb);
    cudaEvent_t e;
    cudaEventCreate(&e);
    const bool holder_alive = with_holder && cudaStreamQuery(sa) == cudaErrorNotReady;
    cudaDeviceSynchronize();
    unsigned long long h[3];
    cudaMemcpy(h, stale, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s resident CTA of another kernel%s: stale words over %d write->read rounds: ld.global.ca %llu, ld.global.nc %llu, ld.global.cg %llu\n",
           with_holder ? "WITH a" : "without a", with_holder ? (holder_alive ? " (still running at the end)" : " (ended early!)") : "",
           iters, h[0], h[1], h[2]);
  }
  cudaError_t e = cudaGetLastError();
  printf("cuda status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
