// Micro-probe: TMA (cp.async.bulk.tensor) load / store throughput per SM as a function of the box
// geometry - row width in bytes, rows per box, rows that are partly out of bounds (zero filled) -
// against a plain 1-D bulk copy of the same size.  All SMs run at once, each over its own slice
// of a buffer that fits L2 (second pass) so that the figures are about the copy engine, not HBM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../voice-tts_b200/csrc/common.cuh"
using namespace bvg;

struct Case {
  const char* name;
  int es;          // element bytes
  int dim0;        // tensor row length in elements (global rows are dim0*es bytes, contiguous rows)
  int box0, box1;  // box: elements per row, rows
  int swz;         // swizzle bytes
  int store;       // 0 load, 1 store
  int bulk1d;      // 1: cp.async.bulk 1-D of box0*box1*es bytes instead
};

#ifndef NBUF
#define NBUF 4
#endif
#ifndef NISS
#define NISS 1
#endif
constexpr int NOPS = 256;

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmap, Case c, const unsigned char* gbase,
                                                long long rows_per_sm, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_all[NISS * NBUF];
  const int box_bytes = c.box0 * c.box1 * c.es;
  const int buf_stride = (box_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NISS * NBUF; ++i) mbar_init(&full_all[i], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < NISS * NBUF * buf_stride / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x % 32 == 0 && threadIdx.x / 32 < NISS) {
    const int iss = threadIdx.x / 32;
    uint64_t* full = full_all + iss * NBUF;
    smem += iss * NBUF * buf_stride;
    const long long row0 = (long long)blockIdx.x * rows_per_sm + iss * (rows_per_sm / NISS / c.box1 * c.box1);
    const int nbox = (int)(rows_per_sm / NISS / c.box1);
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t ph[NBUF] = {};
      const long long t0 = clock64();
      for (int op = 0; op < NOPS; ++op) {
        const int b = op % NBUF;
        const long long row = row0 + (long long)(op % nbox) * c.box1;
        if (c.store) {
          if (op >= NBUF) bulk_wait_group_read<NBUF - 1>();
          tma_store_3d(&tmap, smem + b * buf_stride, 0, (int)row, 0);
          bulk_commit_group();
        } else {
          if (op >= NBUF) { mbar_wait(&full[b], ph[b]); ph[b] ^= 1; }
          mbar_expect_tx(&full[b], (uint32_t)box_bytes);
          if (c.bulk1d) bulk_g2s(smem + b * buf_stride, gbase + row * (long long)c.dim0 * c.es, (uint32_t)box_bytes, &full[b]);
          else tma_load_3d(smem + b * buf_stride, &tmap, 0, (int)row, 0, &full[b]);
        }
      }
      if (c.store) bulk_wait_group<0>();
      else for (int b = 0; b < NBUF; ++b) { mbar_wait(&full[b], ph[b]); ph[b] ^= 1; }
      const long long t1 = clock64();
      if (iss == 0) out[blockIdx.x * 2 + pass] = t1 - t0;
      // re-arm parity bookkeeping for the second pass: barriers have completed an equal number of phases
      // per buffer (NOPS / NBUF + ...): recompute from scratch
      if (pass == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const size_t gbytes = 64ull << 20;
  unsigned char* g;
  cudaMalloc(&g, gbytes);
  cudaMemset(g, 1, gbytes);
  long long* out;
  cudaMalloc(&out, sms * 2 * sizeof(long long));
  std::vector<long long> h(sms * 2);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Case cases[] = {
      {"load bf16 row128B x128 SW128 (full rows)", 2, 64, 64, 128, 128, 0, 0},
      {"load bf16 row128B x256 SW128 (full rows)", 2, 64, 64, 256, 128, 0, 0},
      {"load bf16 row128B x32  SW128 (full rows)", 2, 64, 64, 32, 128, 0, 0},
      {"load bf16 box128B x128 SW128, rows 64B valid", 2, 32, 64, 128, 128, 0, 0},
      {"load bf16 box128B x32  SW128, rows 64B valid", 2, 32, 64, 32, 128, 0, 0},
      {"load bf16 box128B x128 SW128, rows 96B valid", 2, 48, 64, 128, 128, 0, 0},
      {"load bf16 row64B  x128 SW64", 2, 32, 32, 128, 64, 0, 0},
      {"load bf16 row64B  x256 SW64", 2, 32, 32, 256, 64, 0, 0},
      {"load bf16 row32B  x256 SW32", 2, 16, 16, 256, 32, 0, 0},
      {"load bf16 row128B of 1536B rows x128 SW128", 2, 768, 64, 128, 128, 0, 0},
      {"load fp32 row128B x128 noswz (C=32)", 4, 32, 32, 128, 0, 0, 0},
      {"load fp32 row512B x32 noswz (C=128)", 4, 128, 128, 32, 0, 0, 0},
      {"load fp32 row192B x64 noswz (C=48)", 4, 48, 48, 64, 0, 0, 0},
      {"load 1-D bulk 16 KB", 2, 64, 64, 128, 0, 0, 1},
      {"load 1-D bulk 4 KB", 2, 64, 64, 32, 0, 0, 1},
      {"load 1-D bulk 32 KB", 2, 64, 64, 256, 0, 0, 1},
      {"store fp32 row128B x128 noswz (C=32)", 4, 32, 32, 128, 0, 1, 0},
      {"store bf16 row64B x128 noswz (C=32)", 2, 32, 32, 128, 0, 1, 0},
      {"store bf16 row96B x64 noswz (C=48)", 2, 48, 48, 64, 0, 1, 0},
      {"store fp32 row512B x32 noswz (C=128)", 4, 128, 128, 32, 0, 1, 0},
      {"store bf16 row256B x32 noswz (C=128)", 2, 128, 128, 32, 0, 1, 0},
      {"store fp32 row512B of 3072B rows x32", 4, 768, 128, 32, 0, 1, 0},
  };
  printf("%-50s | %9s %9s %9s %9s\n", "case (L2-resident pass)", "clk/op", "clk/row", "B/clk/SM", "GB/s chip");
  for (const Case& c : cases) {
    const long long row_bytes = (long long)c.dim0 * c.es;
    const long long total_rows = (long long)(gbytes / row_bytes);
    long long rows_per_sm = total_rows / sms / c.box1 * c.box1;
    if (rows_per_sm > (long long)c.box1 * 64) rows_per_sm = (long long)c.box1 * 64;   // keep the working set in L2
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)c.dim0, (cuuint64_t)total_rows, 1};
    cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)row_bytes * total_rows};
    cuuint32_t box[3] = {(cuuint32_t)c.box0, (cuuint32_t)c.box1, 1};
    cuuint32_t es3[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = c.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : c.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : c.swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(&m, c.es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, strides,
                     box, es3, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-50s | encode failed %d\n", c.name, (int)r); continue; }
    probe<<<sms, 128, smem>>>(m, c, g, rows_per_sm, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-50s | CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), out, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[2 * i + 1];
    avg /= sms;
    const double clk_op = avg / (NOPS * NISS);
    const int valid0 = c.dim0 < c.box0 ? c.dim0 : c.box0;
    const double bytes = (double)valid0 * c.box1 * c.es;
    printf("%-50s | %9.0f %9.2f %9.1f %9.0f\n", c.name, clk_op, clk_op / c.box1, bytes / clk_op, bytes / clk_op * sms * 1.9);
  }
  return 0;
}
