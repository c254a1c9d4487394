// Internal (non-ABI) interfaces between the translation units of libbvg_b200.
#pragma once
#include "common.cuh"

namespace bvg {

// One dense layer on channels-last data (see conv_simt.cu for the formula).
// A ConvTranspose1d is expressed as a 3-tap conv over the input time axis that
// writes u*Cout_p "phase channels" per input sample (layout.cu, pack_convtr_kernel).
struct ConvArgs {
  const void* in;      // [B, T, Cin_p]
  const void* w;       // Wp[k][Cout_r][Cin_p]
  const float* bias;   // [Cout_r] fp32 (zero in pad rows) or nullptr
  int64_t bias_bs = 0; // elements between the bias rows of consecutive utterances (0: one bias for all; > 0: the per-utterance
                       // bias of the speaker-conditioned generator, indextts/BigVGAN/models.py:224-234 `x + cond(speaker_embedding)`)
  void* out;           // [B, T, out_ld]; channels [0, Cout_n) are written
  const float* res;    // optional fp32 [B, T, out_ld]
  const float* accum;  // optional fp32 [B, T, out_ld]
  float scale;
  int in_dtype, w_dtype, out_dtype;
  int B;
  int64_t T;
  int Cin_p;    // multiple of 16
  int Cout_n;   // channels written per row (multiple of 8)
  int Cout_r;   // rows of Wp per tap (multiple of 128)
  int out_ld;   // row pitch of out/res/accum in elements
  int k, dil;
  int own_sm = 1;   // tcgen05 kernels: 1 = request the SM's whole shared-memory carve-out (no co-resident blocks), 0 = only what is used
};

int conv_simt_launch(const ConvArgs& a, cudaStream_t st);
// tcgen05 implicit-GEMM path: in/w must be bf16.  `variant` selects debug variants (0 = default).
int conv_umma_launch(const ConvArgs& a, int variant, cudaStream_t st);
bool conv_umma_supported(const ConvArgs& a);
// Conv1d + bias followed by Activation1d (alpha/beta log-scale per output channel), bf16 result, one kernel
bool conv_act_fused_supported(const ConvArgs& a);
// with a.res: out = bf16(act(conv + bias + res)), y_out = conv + bias + res (fp32 residual stream; must not alias a.res)
int conv_act_fused_launch(const ConvArgs& a, const float* alpha_log, const float* beta_log, const Taps& taps, cudaStream_t st,
                          float* y_out = nullptr);

// C channels per row are processed; ld (0 = C) is the row pitch in elements
int act1d_cl_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log, const Taps& taps,
                    int B, int64_t T, int C, int in_dtype, int out_dtype, bool fast, cudaStream_t st, int ld = 0);
// One AMPBlock1 unit y = x + c2(a2(c1(a1(x)))) (bigvgan.py:132-141) as one kernel for <= 96-channel stages (amp_unit.cu).
// x / out / accum: channels-last fp32 [B, T, ld] (out bf16 when out_bf16); weights packed bf16 Wp[k][128][Cp] as the
// tcgen05 conv kernels take them (replicated rows for <= 64 channels), bias [128] fp32, alpha / beta [Cp] log scale.
//   out = (conv2(...) + bias2 + x) * scale [+ accum]
struct AmpUnitArgs {
  const float* x = nullptr;
  void* out = nullptr;
  const float* accum = nullptr;
  float scale = 1.f;
  int out_bf16 = 0;
  const void *w1 = nullptr, *w2 = nullptr;
  const float *bias1 = nullptr, *bias2 = nullptr;
  const float *al1 = nullptr, *be1 = nullptr, *al2 = nullptr, *be2 = nullptr;
  Taps taps1, taps2;
  int B = 0;
  int64_t T = 0;
  int C = 0, Cp = 0, ld = 0;
  int k = 0, dil = 1;
};
bool amp_unit_supported(const AmpUnitArgs& a);
int amp_unit_launch(const AmpUnitArgs& a, cudaStream_t st);

int act1d_bct_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log, const Taps& taps,
                     int B, int C, int64_t T, int dtype, bool fast, cudaStream_t st);

int bct_to_btc(void* dst, int out_dtype, const float* src, int B, int C, int Cp, int64_t T, cudaStream_t st);
int btc_to_bct(float* dst, const void* src, int in_dtype, int B, int C, int Cp, int64_t T, cudaStream_t st);
int weight_replica_rows(int Cout_n, int Cout_r);
int pack_conv_weight(void* wp, int dtype, const float* w, int Cout, int Cin, int k, int Cout_r, int Cin_p,
                     cudaStream_t st);
int pack_convtr_weight(void* wp, int dtype, const float* w, int Cin, int Cout, int u, int k, int Cout_p, int Cout_r,
                       int Cin_p, cudaStream_t st);
// k - u even, 0 <= (k - u)/2 <= u: the layer is a 3-tap conv over the input rows (T_out = u * T_in)
static inline bool convtr_shape_ok(int k, int u) { return u >= 1 && k >= u && (k - u) % 2 == 0 && (k - u) / 2 <= u && k <= 2 * u + (k - u) / 2; }
// per-utterance bias rows: out[b][r*Cp + c] = bias[c] + cb[c] + sum_e Wc[c][e] * emb[b][e]  (r < rep; pad entries zero)
int cond_bias_launch(float* out, int64_t out_bs, const float* bias, const float* Wc, const float* cb, const float* emb, int B,
                     int E, int C, int Cp, int rep, cudaStream_t st);
// x = b0 + b1 + b2 (three bf16 terms, exact to fp32's 24 bits): bf16 outputs for activations, fp32 containers for weights
int split3_bf16(void* o0, void* o1, void* o2, const float* src, int64_t n, cudaStream_t st);
int split3_f32(float* o0, float* o1, float* o2, const float* src, int64_t n, cudaStream_t st);
// [B, T, C] fp32 -> [B, T, Cp] (cast to out_dtype, zero pad channels)
int btc_pad_cast(void* dst, int out_dtype, const float* src, int64_t rows, int C, int Cp, cudaStream_t st);
int conv_post_launch(void* dst, int out_i16, const void* src, int in_dtype, const float* w, float bias, int B,
                     int Cp, int64_t T, int use_tanh, cudaStream_t st);
int f32_to_i16(int16_t* dst, const float* src, int64_t n, cudaStream_t st);

// channel padding of the channels-last tensors (the tcgen05 kernels accept multiples of 8 - the zero fill of their
// 64-channel TMA boxes completes the last K step - but 16 keeps every bf16 row a multiple of 32 bytes)
static inline int pad_channels(int c, int gran = 16) { return c < 16 ? 16 : round_up(c, gran); }
static inline size_t dtype_size(int dt) { return dt == BVG_BF16 ? 2 : 4; }

}  // namespace bvg
