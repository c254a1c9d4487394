// Layout conversion and weight packing kernels.
//
// Internal activation layout: channels-last [B, T, Cp], Cp = channels padded to
// a multiple of 16 (pad channels are always zero).  Internal conv weights:
// Wp[tap][Cout_r][Cin_p] (input channel fastest), Cout_r = Cout padded to a
// multiple of 128 - this is directly the K-major A operand of the tcgen05 kernel
// and is also what the fp32 SIMT kernel reads.
#include "conv.cuh"

namespace bvg {

// [B, C, T] (time fastest) -> [B, T, Cp] (channel fastest), zero pad channels
template <typename Tin, typename Tout>
__global__ void bct_to_btc_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, int C, int Cp, int64_t T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const Tin* s = src + (int64_t)b * C * T;
  Tout* d = dst + (int64_t)b * T * Cp;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? ld_mut_f32<Tin>(s + (int64_t)c * T + t) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t t = t0 + i;
    const int c = c0 + threadIdx.x;
    if (t < T && c < Cp) d[t * Cp + c] = from_f32<Tout>(tile[threadIdx.x][i]);
  }
}

// [B, T, Cp] -> [B, C, T]
template <typename Tin, typename Tout>
__global__ void btc_to_bct_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, int C, int Cp, int64_t T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const Tin* s = src + (int64_t)b * T * Cp;
  Tout* d = dst + (int64_t)b * C * T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t t = t0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < Cp) ? ld_mut_f32<Tin>(s + t * Cp + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t t = t0 + threadIdx.x;
    if (c < C && t < T) d[(int64_t)c * T + t] = from_f32<Tout>(tile[threadIdx.x][i]);
  }
}

template <typename Tin, typename Tout>
static int launch_bct_to_btc(void* dst, const void* src, int B, int C, int Cp, int64_t T, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(Cp, 32), (unsigned)B), block(32, 8);
  bct_to_btc_kernel<Tin, Tout><<<grid, block, 0, st>>>((Tout*)dst, (const Tin*)src, C, Cp, T);
  BVG_LAUNCHED();
  return BVG_OK;
}
template <typename Tin, typename Tout>
static int launch_btc_to_bct(void* dst, const void* src, int B, int C, int Cp, int64_t T, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(Cp, 32), (unsigned)B), block(32, 8);
  btc_to_bct_kernel<Tin, Tout><<<grid, block, 0, st>>>((Tout*)dst, (const Tin*)src, C, Cp, T);
  BVG_LAUNCHED();
  return BVG_OK;
}

int bct_to_btc(void* dst, int out_dtype, const float* src, int B, int C, int Cp, int64_t T, cudaStream_t st) {
  if (B <= 0 || T <= 0) return BVG_OK;
  return out_dtype == BVG_BF16 ? launch_bct_to_btc<float, __nv_bfloat16>(dst, src, B, C, Cp, T, st)
                               : launch_bct_to_btc<float, float>(dst, src, B, C, Cp, T, st);
}
int btc_to_bct(float* dst, const void* src, int in_dtype, int B, int C, int Cp, int64_t T, cudaStream_t st) {
  if (B <= 0 || T <= 0) return BVG_OK;
  return in_dtype == BVG_BF16 ? launch_btc_to_bct<__nv_bfloat16, float>(dst, src, B, C, Cp, T, st)
                              : launch_btc_to_bct<float, float>(dst, src, B, C, Cp, T, st);
}

// ---- weight packing -------------------------------------------------------------
// Conv1d weight [Cout, Cin, k] (torch) -> Wp[k][Cout_r][Cin_p]
template <typename Tout>
__global__ void pack_conv_kernel(Tout* __restrict__ wp, const float* __restrict__ w, int Cout, int Cin, int k,
                                 int Cout_r, int Cin_p, int rep_lr) {
  const int64_t n = (int64_t)k * Cout_r * Cin_p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    int co = (int)((i / Cin_p) % Cout_r);
    if (rep_lr) co %= rep_lr;   // narrow layers: the pad rows of the 128-row tile hold replicas (see weight_replica_rows)
    const int j = (int)(i / ((int64_t)Cin_p * Cout_r));
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[((int64_t)co * Cin + ci) * k + j];
    wp[i] = from_f32<Tout>(v);
  }
}

// ConvTranspose1d weight [Cin, Cout, k] (stride u, padding p = (k - u)/2, so T_out = u * T_in) -> a 3-tap conv over
// the INPUT time axis producing u*Cout_p "phase channels" per input sample:
//   out[u*m + r, co] = sum_{tap in 0..2} Wp[tap][r*Cout_p + co][:] . x[m + tap - 1, :]
// torch: out[t] = sum_{m', j : u*m' + j - p = t} W[:, :, j]^T x[m'], i.e. for t = u*m + r and input row m' = m + tap - 1
// the weight index is j = r + p + u*(1 - tap) when it lies in [0, k).  k = 2u, p = u/2 is the BigVGAN v2 plan
// (bigvgan.py:306-312; SURVEY.md 8(a) polyphase form); k = u, p = 0 (taps 0 and 2 all zero) appears in the v1
// generator's plan (indextts/BigVGAN/models.py:154-161 with the published upsample_kernel_sizes).
template <typename Tout>
__global__ void pack_convtr_kernel(Tout* __restrict__ wp, const float* __restrict__ w, int Cin, int Cout, int u, int k,
                                   int Cout_p, int Cout_r, int Cin_p, int rep_lr) {
  const int pad = (k - u) / 2;
  const int64_t n = (int64_t)3 * Cout_r * Cin_p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    int vc = (int)((i / Cin_p) % Cout_r);  // virtual output channel r*Cout_p + co
    if (rep_lr) vc %= rep_lr;
    const int tap = (int)(i / ((int64_t)Cin_p * Cout_r));
    float v = 0.f;
    const int r = vc / Cout_p, co = vc % Cout_p;
    if (ci < Cin && r < u && co < Cout) {
      const int widx = r + pad + u * (1 - tap);
      if (widx >= 0 && widx < k) v = w[((int64_t)ci * Cout + co) * k + widx];
    }
    wp[i] = from_f32<Tout>(v);
  }
}

// Layers whose whole output (Cout_n channels per row) fits 64 (32) MMA rows keep 2 (4) copies of
// their weight rows in the 128-row tile: the tcgen05 kernel then finds the same result in every
// TMEM lane group and all epilogue warps share the time columns (conv_umma2.cu).  Returns the
// replica period in rows (0 = no replication).
int weight_replica_rows(int Cout_n, int Cout_r) {
  if (Cout_r != 128 || Cout_n > 64) return 0;
  return Cout_n <= 32 ? 32 : 64;
}

int pack_conv_weight(void* wp, int dtype, const float* w, int Cout, int Cin, int k, int Cout_r, int Cin_p,
                     cudaStream_t st) {
  const int rep_lr = weight_replica_rows(Cout < 16 ? 16 : round_up(Cout, 16), Cout_r);
  const int64_t n = (int64_t)k * Cout_r * Cin_p;
  const int blocks = (int)(ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096);
  if (dtype == BVG_BF16)
    pack_conv_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)wp, w, Cout, Cin, k, Cout_r, Cin_p, rep_lr);
  else
    pack_conv_kernel<float><<<blocks, 256, 0, st>>>((float*)wp, w, Cout, Cin, k, Cout_r, Cin_p, rep_lr);
  BVG_LAUNCHED();
  return BVG_OK;
}
int pack_convtr_weight(void* wp, int dtype, const float* w, int Cin, int Cout, int u, int k, int Cout_p, int Cout_r,
                       int Cin_p, cudaStream_t st) {
  if (!convtr_shape_ok(k, u)) BVG_FAIL(BVG_EINVAL, "ConvTranspose1d: kernel %d / stride %d is not a 3-tap polyphase layer", k, u);
  const int rep_lr = weight_replica_rows(u * Cout_p, Cout_r);
  const int64_t n = (int64_t)3 * Cout_r * Cin_p;
  const int blocks = (int)(ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096);
  if (dtype == BVG_BF16)
    pack_convtr_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)wp, w, Cin, Cout, u, k, Cout_p, Cout_r,
                                                              Cin_p, rep_lr);
  else
    pack_convtr_kernel<float><<<blocks, 256, 0, st>>>((float*)wp, w, Cin, Cout, u, k, Cout_p, Cout_r, Cin_p, rep_lr);
  BVG_LAUNCHED();
  return BVG_OK;
}

// ---- conv_post (C -> 1, k = 7, pad 3) + clamp/tanh (+ optional int16) ------------------
// reference: bigvgan.py:379-384 and infer_v2.py:740.  in: [B, T, Cp]; w: [7][Cp] fp32 (zero padded);
// out: [B, 1, T] fp32 (or int16).  Memory-bound (reads Cp channels, writes 1).
template <typename Tin, typename Tout>
__global__ void conv_post_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, const float* __restrict__ w,
                                 float bias, int Cp, int64_t T, int use_tanh) {
  extern __shared__ float sw[];  // [7][Cp]
  for (int i = threadIdx.x; i < 7 * Cp; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const Tin* s = src + (int64_t)b * T * Cp;
  float acc = bias;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int64_t tt = t + j - 3;
    if (tt < 0 || tt >= T) continue;
    const Tin* row = s + tt * Cp;
    for (int c = 0; c < Cp; c += 8) {
      float v[8];
      if (sizeof(Tin) == 2) {
        const uint4 raw = BVG_LDG(reinterpret_cast<const uint4*>(row + c));
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[2 * q] = __uint_as_float(r[q] << 16);
          v[2 * q + 1] = __uint_as_float(r[q] & 0xffff0000u);
        }
      } else {
        const float4 a = BVG_LDG(reinterpret_cast<const float4*>(row + c));
        const float4 bq = BVG_LDG(reinterpret_cast<const float4*>(row + c) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(sw[j * Cp + c + q], v[q], acc);
    }
  }
  acc = use_tanh ? tanhf(acc) : fminf(fmaxf(acc, -1.0f), 1.0f);
  if (sizeof(Tout) == 2) {
    // clamp(32767 * wav, -32767, 32767) -> int16 (infer_v2.py:740; torch .type(int16) truncates)
    float q = fminf(fmaxf(32767.0f * acc, -32767.0f), 32767.0f);
    reinterpret_cast<int16_t*>(dst)[(int64_t)b * T + t] = (int16_t)q;
  } else {
    reinterpret_cast<float*>(dst)[(int64_t)b * T + t] = acc;
  }
}

// Tiled variant: the 256 + 6 input rows of a block are staged in shared memory once (coalesced 16-byte ld.global.cg), so
// every row is fetched from L2 once instead of 7 times; rows are padded by 16 bytes, which makes the per-thread 16-byte
// reads (thread t reads row t + j) conflict-free for 64- and 128-byte rows.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
conv_post_tiled_kernel(Tout* __restrict__ dst, const Tin* __restrict__ src, const float* __restrict__ w, float bias, int Cp,
                       int64_t T, int use_tanh) {
  extern __shared__ __align__(16) unsigned char cp_smem[];
  float* sw = reinterpret_cast<float*>(cp_smem);                 // [7][Cp]
  const int row_bytes = Cp * (int)sizeof(Tin);
  const int pitch = row_bytes + 16;
  unsigned char* tile = cp_smem + ((7 * Cp * 4 + 15) & ~15);     // [262][pitch]
  for (int i = threadIdx.x; i < 7 * Cp; i += 256) sw[i] = w[i];
  const int b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * 256;
  const unsigned char* s = reinterpret_cast<const unsigned char*>(src) + (int64_t)b * T * row_bytes;
  const int vec_per_row = row_bytes / 16;
  for (int i = threadIdx.x; i < 262 * vec_per_row; i += 256) {
    const int r = i / vec_per_row, v = i % vec_per_row;
    const int64_t tt = t0 - 3 + r;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);                      // rows outside [0, T): the conv's zero padding
    if (tt >= 0 && tt < T) val = BVG_LDG(reinterpret_cast<const uint4*>(s + tt * row_bytes) + v);
    *reinterpret_cast<uint4*>(tile + r * pitch + v * 16) = val;
  }
  __syncthreads();
  const int64_t t = t0 + threadIdx.x;
  if (t >= T) return;
  float acc = bias;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const unsigned char* row = tile + (threadIdx.x + j) * pitch;
    for (int c = 0; c < Cp; c += 8) {
      float v[8];
      if (sizeof(Tin) == 2) {
        const uint4 raw = *reinterpret_cast<const uint4*>(row + c * 2);
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[2 * q] = __uint_as_float(r[q] << 16);
          v[2 * q + 1] = __uint_as_float(r[q] & 0xffff0000u);
        }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(row + c * 4);
        const float4 bq = *reinterpret_cast<const float4*>(row + c * 4 + 16);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(sw[j * Cp + c + q], v[q], acc);
    }
  }
  acc = use_tanh ? tanhf(acc) : fminf(fmaxf(acc, -1.0f), 1.0f);
  if (sizeof(Tout) == 2) {
    float q = fminf(fmaxf(32767.0f * acc, -32767.0f), 32767.0f);
    reinterpret_cast<int16_t*>(dst)[(int64_t)b * T + t] = (int16_t)q;
  } else {
    reinterpret_cast<float*>(dst)[(int64_t)b * T + t] = acc;
  }
}

template <typename Tin, typename Tout>
static bool conv_post_tiled(void* dst, const void* src, const float* w, float bias, int B, int Cp, int64_t T, int use_tanh,
                            cudaStream_t st) {
  const int smem = ((7 * Cp * 4 + 15) & ~15) + 262 * (Cp * (int)sizeof(Tin) + 16);
  if (smem > 48 * 1024 || Cp % 8 != 0 || (reinterpret_cast<uintptr_t>(src) & 15)) return false;
  dim3 grid((unsigned)ceil_div(T, 256), (unsigned)B);
  conv_post_tiled_kernel<Tin, Tout><<<grid, 256, smem, st>>>((Tout*)dst, (const Tin*)src, w, bias, Cp, T, use_tanh);
  return true;
}

int conv_post_launch(void* dst, int out_i16, const void* src, int in_dtype, const float* w, float bias, int B,
                     int Cp, int64_t T, int use_tanh, cudaStream_t st) {
  if (B <= 0 || T <= 0) return BVG_OK;
  {
    bool done;
    if (in_dtype == BVG_BF16)
      done = out_i16 ? conv_post_tiled<__nv_bfloat16, int16_t>(dst, src, w, bias, B, Cp, T, use_tanh, st)
                     : conv_post_tiled<__nv_bfloat16, float>(dst, src, w, bias, B, Cp, T, use_tanh, st);
    else
      done = out_i16 ? conv_post_tiled<float, int16_t>(dst, src, w, bias, B, Cp, T, use_tanh, st)
                     : conv_post_tiled<float, float>(dst, src, w, bias, B, Cp, T, use_tanh, st);
    if (done) { BVG_LAUNCHED(); return BVG_OK; }
  }
  dim3 grid((unsigned)ceil_div(T, 256), (unsigned)B);
  const int smem = 7 * Cp * (int)sizeof(float);
  if (in_dtype == BVG_BF16) {
    if (out_i16)
      conv_post_kernel<__nv_bfloat16, int16_t><<<grid, 256, smem, st>>>((int16_t*)dst, (const __nv_bfloat16*)src, w,
                                                                        bias, Cp, T, use_tanh);
    else
      conv_post_kernel<__nv_bfloat16, float><<<grid, 256, smem, st>>>((float*)dst, (const __nv_bfloat16*)src, w,
                                                                      bias, Cp, T, use_tanh);
  } else {
    if (out_i16)
      conv_post_kernel<float, int16_t><<<grid, 256, smem, st>>>((int16_t*)dst, (const float*)src, w, bias, Cp, T,
                                                                use_tanh);
    else
      conv_post_kernel<float, float><<<grid, 256, smem, st>>>((float*)dst, (const float*)src, w, bias, Cp, T,
                                                              use_tanh);
  }
  BVG_LAUNCHED();
  return BVG_OK;
}

// wav fp32 -> int16 (used when the int16 result is wanted from an fp32 wav buffer)
__global__ void f32_to_i16_kernel(int16_t* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (int16_t)fminf(fmaxf(32767.0f * src[i], -32767.0f), 32767.0f);
}
int f32_to_i16(int16_t* dst, const float* src, int64_t n, cudaStream_t st) {
  if (n <= 0) return BVG_OK;
  const int blocks = (int)(ceil_div(n, 256) < 2048 ? ceil_div(n, 256) : 2048);
  f32_to_i16_kernel<<<blocks, 256, 0, st>>>(dst, src, n);
  BVG_LAUNCHED();
  return BVG_OK;
}

// ---- speaker-conditioned generator (indextts/BigVGAN/models.py:224-234) -------------------------------------------
// `x = conv_pre(x) + cond_layer(e)` and `x = ups[i](x) + conds[i](e)`: the 1x1 convs act on a length-1 sequence, so each is
// a per-utterance vector that joins the layer's bias.  One warp per (utterance, channel): dot over the embedding, written
// to all `rep` phase-channel copies of a ConvTranspose layer.  out rows are pre-zeroed in their pad entries by the caller.
__global__ void cond_bias_kernel(float* __restrict__ out, int64_t out_bs, const float* __restrict__ bias,
                                 const float* __restrict__ Wc, const float* __restrict__ cb, const float* __restrict__ emb,
                                 int E, int C, int Cp, int rep) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (c >= C) return;
  float acc = 0.f;
  for (int e = lane; e < E; e += 32) acc = fmaf(Wc[(int64_t)c * E + e], BVG_LDG(emb + (int64_t)b * E + e), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float v = (bias ? bias[c] : 0.f) + (acc + (cb ? cb[c] : 0.f));
    for (int r = 0; r < rep; ++r) out[(int64_t)b * out_bs + (int64_t)r * Cp + c] = v;
  }
}
int cond_bias_launch(float* out, int64_t out_bs, const float* bias, const float* Wc, const float* cb, const float* emb, int B,
                     int E, int C, int Cp, int rep, cudaStream_t st) {
  if (B <= 0 || C <= 0) return BVG_OK;
  dim3 grid((unsigned)ceil_div(C, 8), (unsigned)B);
  cond_bias_kernel<<<grid, 256, 0, st>>>(out, out_bs, bias, Wc, cb, emb, E, C, Cp, rep);
  BVG_LAUNCHED();
  return BVG_OK;
}

// GPT latent [B, T, C] fp32 (models.py:220 `x.transpose(1, 2)`: already channels-last) -> conv_pre operand [B, T, Cp]
template <typename Tout>
__global__ void btc_pad_cast_kernel(Tout* __restrict__ dst, const float* __restrict__ src, int64_t rows, int C, int Cp) {
  const int64_t n = rows * Cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int64_t r = i / Cp;
    dst[i] = from_f32<Tout>(c < C ? src[r * C + c] : 0.f);
  }
}
int btc_pad_cast(void* dst, int out_dtype, const float* src, int64_t rows, int C, int Cp, cudaStream_t st) {
  if (rows <= 0) return BVG_OK;
  const int64_t n = rows * Cp;
  const int blocks = (int)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
  if (out_dtype == BVG_BF16) btc_pad_cast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)dst, src, rows, C, Cp);
  else btc_pad_cast_kernel<float><<<blocks, 256, 0, st>>>((float*)dst, src, rows, C, Cp);
  BVG_LAUNCHED();
  return BVG_OK;
}

// ---- fp32 values as sums of three bf16 terms (fp32-accurate convolutions on the bf16 tensor cores) -----------------------
// x = b0 + b1 + b2 with b0 = bf16(x), b1 = bf16(x - b0), b2 = bf16(x - b0 - b1): the subtractions are exact in fp32 and
// three 8-bit mantissas cover fp32's 24 bits (bf16 has fp32's exponent range, so nothing underflows that fp32 keeps).
__global__ void split3_bf16_kernel(__nv_bfloat16* __restrict__ o0, __nv_bfloat16* __restrict__ o1, __nv_bfloat16* __restrict__ o2,
                                   const float* __restrict__ src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = BVG_LDG(src + i);
    const __nv_bfloat16 b0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b0);
    const __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b1);
    o0[i] = b0; o1[i] = b1; o2[i] = __float2bfloat16_rn(r2);
  }
}
int split3_bf16(void* o0, void* o1, void* o2, const float* src, int64_t n, cudaStream_t st) {
  if (n <= 0) return BVG_OK;
  const int blocks = (int)(ceil_div(n, 256) < 148 * 32 ? ceil_div(n, 256) : 148 * 32);
  split3_bf16_kernel<<<blocks, 256, 0, st>>>((__nv_bfloat16*)o0, (__nv_bfloat16*)o1, (__nv_bfloat16*)o2, src, n);
  BVG_LAUNCHED();
  return BVG_OK;
}
// the same split kept in fp32 containers (weights: each term then goes through the ordinary bf16 packers)
__global__ void split3_f32_kernel(float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2,
                                  const float* __restrict__ src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = src[i];
    const float b0 = __bfloat162float(__float2bfloat16_rn(x));
    const float r1 = x - b0;
    const float b1 = __bfloat162float(__float2bfloat16_rn(r1));
    o0[i] = b0; o1[i] = b1; o2[i] = __bfloat162float(__float2bfloat16_rn(r1 - b1));
  }
}
int split3_f32(float* o0, float* o1, float* o2, const float* src, int64_t n, cudaStream_t st) {
  if (n <= 0) return BVG_OK;
  const int blocks = (int)(ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096);
  split3_f32_kernel<<<blocks, 256, 0, st>>>(o0, o1, o2, src, n);
  BVG_LAUNCHED();
  return BVG_OK;
}


}  // namespace bvg
