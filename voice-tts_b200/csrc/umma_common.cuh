// Pieces shared by the two tcgen05 conv kernels (channel-major conv_umma.cu, time-major conv_umma_t.cu).
#pragma once
#include <cuda.h>

#include "conv.cuh"

namespace bvg {

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int row_bytes, int base_mode) {
  // K-major, swizzled: SBO = 8 rows; LBO unused (1); version 1 (sm_100)
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (base_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}


int make_map_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                int row_bytes);
int make_map_4d_w(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                  uint32_t b2, int row_bytes);
int make_map_any(CUtensorMap* m, const void* base, int es, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes);
int umma_sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per device and kernel instantiation instead of once per launch
// (`done`: one bit per device ordinal; the attribute is per device)
template <typename K>
static inline int smem_attr_once(K kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  BVG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return BVG_OK;
  BVG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  return BVG_OK;
}
// v2 channel-major kernel (conv_umma2.cu): 128-byte operand rows, TMA-store epilogue
bool conv_umma2_supported(const ConvArgs& a);
int conv_umma2_launch(const ConvArgs& a, int variant, cudaStream_t st);
// v2 kernel with the following Activation1d fused into its epilogue (conv_umma2a.cu); bf16 output only
bool conv_umma2a_supported(const ConvArgs& a);
// a.res != nullptr: out = act(conv + bias + res) and y_out = conv + bias + res (fp32, same layout; must not alias a.res)
int conv_umma2a_launch(const ConvArgs& a, const float* alpha_log, const float* beta_log, const Taps& taps, cudaStream_t st,
                       float* y_out = nullptr);
int conv_umma_t_launch(const ConvArgs& a, int variant, cudaStream_t st);
bool conv_umma_t_fits(const ConvArgs& a);

}  // namespace bvg
