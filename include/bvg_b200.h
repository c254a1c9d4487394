/*
 * bvg_b200.h — C ABI of libbvg_b200.so: the BigVGAN v2 vocoder hot path of
 * caishiqing/voice-tts (IndexTTS2), hand-written CUDA for sm_100a (B200).
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain
 * pointers / sizes / a CUDA stream handle (no torch or C++ types), returns 0 on
 * success or a negative BVG_E* code, and never throws.  `bvg_last_error()`
 * returns a thread-local message for the last failure.  All buffers are owned
 * by the caller unless stated; the library only owns what `bvg_create`
 * allocates (packed weights + workspace) and frees it in `bvg_destroy`.
 * There is no CPU fallback: with no sm_100 device every compute entry point
 * fails with BVG_ENODEV.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   bvg_act1d_fwd        <- pybind `anti_alias_activation_cuda.forward(input, up_ftr, down_ftr, alpha, beta)`
 *                           indextts/s2mel/modules/bigvgan/alias_free_activation/cuda/anti_alias_activation.cpp:19-23
 *                           -> `fwd_cuda` anti_alias_activation_cuda.cu:212-246, called from
 *                           cuda/activation1d.py:21-27 (FusedAntiAliasActivation.forward); semantics of
 *                           torch/act.py:25-30 (replicate-pad edges exactly as the torch path)
 *   bvg_conv1d_fwd       <- torch.nn.Conv1d as built at bigvgan.py:59-66,76-83,285-287,348-350
 *   bvg_convtr1d_fwd     <- torch.nn.ConvTranspose1d as built at bigvgan.py:306-312
 *   bvg_amp_unit_fwd     <- one iteration of AMPBlock1.forward, bigvgan.py:132-141 (+ the mean of :369-375)
 *   bvg_create/..._fwd   <- BigVGAN.__init__/forward/remove_weight_norm  bigvgan.py:266-400,
 *                           called from indextts/infer_v2.py:155-158,735
 *   bvg_vocoder_fwd_cond <- the speaker-conditioned IndexTTS-v1 generator, indextts/BigVGAN/models.py:130-250
 *                           (`forward(x, mel_ref, lens)` :212-250 minus its ECAPA speaker encoder), called from
 *                           indextts/infer.py:476,646 as `wav, _ = self.bigvgan(latent, auto_conditioning.transpose(1, 2))`
 *   bvg_vocoder_fwd_host <- the host round trip of indextts/infer_v2.py:735-744 (mel on host -> wav on host,
 *                           optional int16 quantisation `clamp(32767*wav)` of :740)
 *   bvg_s2mel_tail_*     <- what DiT.forward does after its transformer, indextts/s2mel/modules/diffusion_transformer.py:245-256
 *                           (conv1, WN of wavenet.py:103-164, res_projection, FinalLayer :82-99, conv2)
 *   bvg_cfm_euler_step   <- the per-step arithmetic of BASECFM.solve_euler, indextts/s2mel/modules/flow_matching.py:103-112
 */
#ifndef BVG_B200_H_
#define BVG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVG_ABI_VERSION 2

/* status codes */
#define BVG_OK 0
#define BVG_EINVAL (-1)   /* bad shape / argument / NULL pointer          */
#define BVG_EDTYPE (-2)   /* unsupported dtype                            */
#define BVG_EALIGN (-3)   /* pointer not aligned as documented            */
#define BVG_ECUDA (-4)    /* CUDA runtime/driver error (see last_error)   */
#define BVG_ENODEV (-5)   /* no sm_100 device / kernel image not loadable */
#define BVG_ENOMEM (-6)   /* device or host allocation failed             */
#define BVG_ESTATE (-7)   /* handle not finalized / weight missing        */

/* element types of activations at the ABI */
#define BVG_F32 0
#define BVG_BF16 1
#define BVG_F16 2 /* bvg_act1d_fwd only: the reference kernel dispatches half too (type_shim.h:20-43) */

/* precision modes of the whole-vocoder handle */
#define BVG_MODE_FP32 0 /* fp32 storage, fp32 SIMT convs, accurate sin: <=1e-5 rel. vs fp32 reference */
#define BVG_MODE_BF16 1 /* bf16 conv operands on tcgen05 tensor cores, fp32 accumulate + fp32 residual  */

/* snake kinds */
#define BVG_SNAKE 0     /* x + 1/(a+1e-9) sin^2(a x)   (activations.py:46-59)   */
#define BVG_SNAKEBETA 1 /* x + 1/(b+1e-9) sin^2(a x)   (activations.py:107-120) */

typedef void* bvg_stream_t; /* a cudaStream_t (CUstream); NULL = legacy default stream */
typedef struct bvg_vocoder bvg_vocoder;

int bvg_abi_version(void);
const char* bvg_last_error(void);
/* number of kernels this library has launched in the calling process so far */
uint64_t bvg_launch_count(void);

/* ------------------------------------------------------------------------
 * Fused anti-aliased activation: up x2 (12-tap kaiser-sinc, replicate pad 5|5)
 * -> Snake/SnakeBeta -> down x2 (12-tap, replicate pad 5|6), one kernel.
 *   dst, src : [B, C, T] contiguous, time fastest (the reference layout), dtype `dtype` (BVG_F32 / BVG_BF16 / BVG_F16; fp32 arithmetic)
 *   alpha_log, beta_log : [C] fp32, LOG scale (the kernel applies exp), device
 *   up_taps, down_taps  : [12] fp32 HOST arrays (Activation1d.upsample.filter / .downsample.lowpass.filter;
 *                         construction-time constants, passed to the kernel as launch parameters)
 *   flags : bit0 = use fast sin (MUFU) instead of the accurate one.
 * T == 0 or B*C == 0 is a no-op that returns BVG_OK (as the reference, .cu:193-196).
 * dst may not alias src.
 */
#define BVG_ACT_FAST_SIN 1
int bvg_act1d_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                  const float* up_taps, const float* down_taps, int B, int C, int64_t T,
                  int dtype, int flags, bvg_stream_t stream);

/* Same operator on channels-last data [B, T, C] (the library's internal layout).
 * in_dtype/out_dtype may differ (fp32 residual stream in, bf16 MMA operand out). */
int bvg_act1d_cl_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                     const float* up_taps, const float* down_taps, int B, int64_t T, int C,
                     int in_dtype, int out_dtype, int flags, bvg_stream_t stream);

/* ------------------------------------------------------------------------
 * Stand-alone dense layers on the reference layout (used by the parity tests
 * and by callers that want single layers): fp32 [B,C,T] in/out, fp32 weights in
 * torch layout.  `mode` selects fp32 SIMT or bf16 tcgen05 arithmetic.
 *   conv1d   : weight [Cout, Cin, k], bias [Cout] or NULL, zero padding (k-1)*dil/2
 *   convtr1d : weight [Cin, Cout, k], padding (k - stride)/2 with k - stride even and <= 2*stride -> T_out = stride*T
 */
int bvg_conv1d_fwd(float* dst, const float* src, const float* weight, const float* bias,
                   int B, int Cin, int Cout, int64_t T, int k, int dilation, int mode,
                   bvg_stream_t stream);
int bvg_convtr1d_fwd(float* dst, const float* src, const float* weight, const float* bias,
                     int B, int Cin, int Cout, int64_t T, int k, int stride, int mode,
                     bvg_stream_t stream);
/* Conv1d with the AMPBlock1 residual and the resblock mean folded into its epilogue
 * (bigvgan.py:132-141 `x = xt + x`, :369-375 `xs += ...; x = xs / num_kernels`):
 *   dst = (conv1d(src) + bias + res) * scale + accum,   res / accum: fp32 [B, Cout, T] or NULL
 * (accum without res is evaluated by the plain kernels).  out_bf16 != 0 rounds the result to bf16
 * before it is returned as fp32 (what the next ConvTranspose1d consumes in BVG_MODE_BF16). */
int bvg_conv1d_res_fwd(float* dst, const float* src, const float* weight, const float* bias, const float* res,
                       const float* accum, float scale, int out_bf16, int B, int Cin, int Cout, int64_t T, int k,
                       int dilation, int mode, bvg_stream_t stream);
/* Conv1d followed by the anti-aliased activation, as `xt = c1(xt); xt = a2(xt)` in AMPBlock1.forward
 * (bigvgan.py:136-138): dst = Activation1d_{alpha,beta}(conv1d(src) + bias).  alpha_log / beta_log: [Cout] fp32
 * log-scale device arrays; up_taps / down_taps: 12 host floats each (as bvg_act1d_fwd).  BVG_MODE_BF16 runs ONE
 * tcgen05 kernel (the fp32 accumulator goes through the activation in registers, the result is rounded to bf16);
 * BVG_MODE_FP32 runs the fp32 conv and the fp32 activation kernel back to back. */
int bvg_conv1d_act_fwd(float* dst, const float* src, const float* weight, const float* bias, const float* alpha_log,
                       const float* beta_log, const float* up_taps, const float* down_taps, int B, int Cin, int Cout,
                       int64_t T, int k, int dilation, int mode, bvg_stream_t stream);
/* Second conv of an AMP unit with its residual AND the first activation of the next unit
 * (bigvgan.py:138-139 `xt = c2(xt); x = xt + x`, then :134 `xt = a1(x)` of the next loop iteration):
 *   dst_y   = conv1d(src) + bias + res                 (fp32 residual stream, [B, Cout, T])
 *   dst_act = Activation1d_{alpha,beta}(dst_y)         (rounded to bf16 in BVG_MODE_BF16)
 * One tcgen05 kernel in BVG_MODE_BF16; conv + activation kernels in BVG_MODE_FP32. */
int bvg_conv1d_res_act_fwd(float* dst_act, float* dst_y, const float* src, const float* weight, const float* bias,
                           const float* res, const float* alpha_log, const float* beta_log, const float* up_taps,
                           const float* down_taps, int B, int Cin, int Cout, int64_t T, int k, int dilation, int mode,
                           bvg_stream_t stream);

/* One whole AMPBlock1 unit - one iteration of the loop in AMPBlock1.forward (bigvgan.py:132-141):
 *   xt = a1(x); xt = c1(xt); xt = a2(xt); xt = c2(xt); x = xt + x
 * with the resblock mean of bigvgan.py:369-375 foldable into the result:
 *   dst = (x + conv1d_k(act2(conv1d_{k,dil}(act1(x)) )) ) * scale + accum        x, dst, accum: fp32 [B, C, T]
 * w1 / w2: [C, C, k] (torch layout), b1 / b2: [C] or NULL, alpha*_log / beta*_log: [C] fp32 log-scale device arrays,
 * up_taps / down_taps: 12 host floats (both activations share them, as every Activation1d of the reference does).
 * BVG_MODE_BF16 runs ONE kernel for C <= 96 (the activations run in registers beside the tcgen05 convolutions, the
 * intermediates never leave the SM); otherwise, or with flags bit 0 set, the four layers run one after the other.
 * flags bit 1: fail with BVG_EINVAL instead of falling back to the layer-by-layer form. */
#define BVG_UNIT_LAYERWISE 1
#define BVG_UNIT_REQUIRE_FUSED 2
int bvg_amp_unit_fwd(float* dst, const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* alpha1_log, const float* beta1_log, const float* alpha2_log, const float* beta2_log,
                     const float* up_taps, const float* down_taps, const float* accum, float scale, int out_bf16,
                     int B, int C, int64_t T, int k, int dilation, int mode, int flags, bvg_stream_t stream);

/* ------------------------------------------------------------------------
 * Whole generator.  Build: bvg_create -> bvg_set_tensor for every state-dict
 * tensor (folded weights, reference key names) -> bvg_finalize -> forward calls.
 */
typedef struct bvg_config {
  int num_mels;                 /* 80 */
  int upsample_initial_channel; /* 1536 */
  int num_upsamples;            /* <= 8 */
  int upsample_rates[8];        /* {4,4,2,2,2,2} */
  int upsample_kernel_sizes[8]; /* {8,8,4,4,4,4}; k - rate even, 0 <= (k - rate)/2 <= rate (padding (k - rate)/2) */
  int num_kernels;              /* <= 4 */
  int resblock_kernel_sizes[4]; /* {3,7,11} */
  int num_dilations;            /* <= 4 */
  int resblock_dilations[4][4]; /* [kernel][layer] {1,3,5} */
  int snake_kind;               /* BVG_SNAKE / BVG_SNAKEBETA */
  int snake_logscale;           /* 1: parameters are log-scale */
  int use_tanh_at_final;
  int use_bias_at_final;
  int mode;                     /* BVG_MODE_* */
  int device;                   /* CUDA device ordinal */
  /* --- ABI 2: the speaker-conditioned v1 generator (indextts/BigVGAN/models.py:130-250); all 0 for BigVGAN v2 --- */
  int input_channels_last;      /* 1: the input is [B, T0, num_mels] (the GPT latent, num_mels = gpt_dim; models.py:220) */
  int cond_dim;                 /* speaker_embedding_dim; > 0 adds cond_layer(e) after conv_pre (models.py:224) */
  int cond_each_up;             /* cond_d_vector_in_each_upsampling_layer: conds[i](e) after ups[i] (models.py:233-234) */
} bvg_config;

int bvg_create(const bvg_config* cfg, bvg_vocoder** out);
void bvg_destroy(bvg_vocoder* v);
/* `name` is a reference state-dict key ("conv_pre.weight", "ups.0.0.bias",
 * "resblocks.3.convs1.0.weight", "resblocks.3.activations.2.act.alpha",
 * "activation_post.upsample.filter", "conv_post.weight", ...); `data` is fp32,
 * contiguous, on the handle's device (is_device=1) or on the host (0); the
 * library copies/packs it before returning.  numel is checked against the config. */
int bvg_set_tensor(bvg_vocoder* v, const char* name, const float* data, int64_t numel, int is_device);
int bvg_finalize(bvg_vocoder* v);
/* bytes of device workspace needed for (B, T0); grows the handle's arena. */
int64_t bvg_workspace_bytes(const bvg_vocoder* v, int B, int T0);

/* mel [B, num_mels, T0] fp32 device -> wav [B, 1, T0*prod(rates)] fp32 device, on `stream`. */
int bvg_vocoder_fwd(bvg_vocoder* v, const float* mel, float* wav, int B, int T0, bvg_stream_t stream);
/* Speaker-conditioned generator (cfg.cond_dim > 0): latent [B, T0, num_mels] fp32 device (cfg.input_channels_last = 1;
 * [B, num_mels, T0] if 0), spk_emb [B, cond_dim] fp32 device (the output of the caller's speaker encoder, models.py:213),
 * wav [B, 1, T0*prod(rates)].  The 1x1 convs `cond_layer` / `conds.{i}` (tensor names "cond_layer.weight|bias",
 * "conds.<i>.weight|bias") are evaluated per utterance and folded into the bias of conv_pre / ups[i]. */
int bvg_vocoder_fwd_cond(bvg_vocoder* v, const float* latent, const float* spk_emb, float* wav, int B, int T0,
                         bvg_stream_t stream);
/* Host buffers: H2D, forward, D2H on `stream`, then waits for completion.
 * wav_dtype: 0 = fp32 wav in [-1,1]; 1 = int16 `clamp(32767*wav, -32767, 32767)` (infer_v2.py:740). */
int bvg_vocoder_fwd_host(bvg_vocoder* v, const float* mel_host, void* wav_host, int wav_dtype,
                         int B, int T0, bvg_stream_t stream);
/* options: "graph" (CUDA-graph replay of the layer sequence: 0 never, 1 always, 2 [default] from the second forward of a (B, T0) shape on), "conv_impl" (0 auto, 1 simt, 2 tcgen05; 3 in BVG_MODE_FP32: every
 * convolution as "split_terms" (3 [default], 6 or 9) bf16 tcgen05 passes over three-term bf16 splits of the fp32 operands -
 * 93 / 96 / 96 dB on the full generator, limited by the tensor cores' fp32 accumulation, 8x / 4.5x / 3x faster than the SIMT kernels),
 * "fast_sin" (0/1), "workspace_mb" (micro-batching cap), "profile" (0/1: one CUDA-event pair per launch, read back with
 * bvg_profile_read), "fuse_act" (0 off, 1 measured policy, 2 always: conv1 + following activation in one kernel),
 * "fuse_res" / "fuse_res_min_kc" (conv2 + residual + next activation in one kernel: 0 off, 1 when k*Cin >= min_kc,
 * 2 always), "fuse_unit" (1: whole AMP units of <= 96-channel stages as ONE kernel, bvg_amp_unit_fwd; 0 [default, faster]: layer by
 * layer), "streams" (3 [default]: the AMP blocks of a stage on separate internal streams that fork from and join the caller's
 * stream; 1: serial; the result is bit-identical either way), "conv_own_sm" (1 [default]: the persistent tcgen05 conv kernels request the whole shared-memory carve-out of their SM;
 * 0: they leave room for one shared-memory-free block of another stream beside them), "pdl" (1: programmatic dependent launch of the
 * conv / activation kernels; 0 [default]: measured no gain), "fold" (1 [default]: resblock convolutions of <= 64-channel stages run
 * as time-folded F*C-channel layers over the same tensors viewed as [T/F, F*C] where that saves tensor-core instructions; 0: off).  Options that change what a forward enqueues drop captured graphs. */
int bvg_set_option(bvg_vocoder* v, const char* key, int value);
/* per-kernel CUDA-event timing (set option "profile"=1 first; disables graph replay while on):
 * category 0 = tcgen05 conv, 1 = SIMT conv, 2 = fused activation, 3 = other, 4 = whole AMP unit in one kernel
 * (work = the flops of its two convolutions).  Returns the summed
 * duration [ms], the summed algorithmic work (flops for convs, bytes for activations) and the
 * number of launches since the last read; reading category 3 clears the records (read it last). */
int bvg_profile_read(bvg_vocoder* v, int category, double* ms, double* work, int* launches);
/* writes one CSV line per recorded launch (category, shape, ms, work, rate) to `path`; does not clear */
int bvg_profile_dump(bvg_vocoder* v, const char* path);
/* introspection for benchmarks: kernels launched by the last forward */
int bvg_last_forward_launches(const bvg_vocoder* v);

/* ------------------------------------------------------------------------
 * The s2mel tail in front of the vocoder: what DiT.forward does after its transformer
 * (indextts/s2mel/modules/diffusion_transformer.py:245-256) - conv1 -> WN (indextts/s2mel/modules/wavenet.py:103-164, reflect-padded
 * SConv1d in_layers, gated tanh * sigmoid, res / skip 1x1 convs, x_mask) conditioned on t_embedder2(t) -> + res_projection(x_res)
 * -> FinalLayer (LayerNorm, adaLN modulate from t1, Linear) -> conv2 - and one Euler / classifier-free-guidance update of
 * BASECFM.solve_euler (indextts/s2mel/modules/flow_matching.py:85-113).  Same build protocol as the vocoder handle:
 * create -> set_tensor (folded weights, reference state-dict names relative to the DiT module: "conv1.weight",
 * "t_embedder2.mlp.0.weight", "t_embedder2.freqs", "wavenet.cond_layer.weight", "wavenet.in_layers.3.bias",
 * "wavenet.res_skip_layers.7.weight", "res_projection.bias", "final_layer.adaLN_modulation.1.weight",
 * "final_layer.linear.weight", "conv2.bias", ... ) -> finalize -> forward calls.
 */
typedef struct bvg_s2mel_config {
  int hidden;        /* wavenet.hidden_dim (512) */
  int dit_hidden;    /* DiT.hidden_dim (512): width of x_res.  t1 = DiT.t_embedder(t) feeds FinalLayer(wavenet.hidden_dim), so the
                        reference itself needs dit_hidden == hidden whenever the wavenet head is used */
  int n_layers;      /* wavenet.num_layers (8) */
  int kernel_size;   /* wavenet.kernel_size (5), odd */
  int dilation_rate; /* wavenet.dilation_rate: 1 (the only value built) */
  int out_channels;  /* DiT.in_channels (80 mel bins) */
  int freq_dim;      /* TimestepEmbedder.frequency_embedding_size (256) */
  int mode;          /* BVG_MODE_* */
  int device;
} bvg_s2mel_config;
typedef struct bvg_s2mel_tail bvg_s2mel_tail;

int bvg_s2mel_tail_create(const bvg_s2mel_config* cfg, bvg_s2mel_tail** out);
void bvg_s2mel_tail_destroy(bvg_s2mel_tail* h);
int bvg_s2mel_tail_set_tensor(bvg_s2mel_tail* h, const char* name, const float* data, int64_t numel, int is_device);
int bvg_s2mel_tail_finalize(bvg_s2mel_tail* h);
int64_t bvg_s2mel_tail_workspace_bytes(const bvg_s2mel_tail* h, int B, int T);
/* x_res [B, T, dit_hidden] fp32 (the transformer output after skip_linear, diffusion_transformer.py:243-244),
 * x_lens [B] int32 or NULL (x_mask = t < x_lens[b]; NULL: all frames valid), t [B] fp32 (the solver's time),
 * t1 [B, hidden] fp32 (DiT.t_embedder(t), computed by the caller for the transformer anyway) -> out [B, out_channels, T] fp32.
 * All device pointers; T > (kernel_size - 1) / 2. */
int bvg_s2mel_tail_fwd(bvg_s2mel_tail* h, const float* x_res, const int* x_lens, const float* t, const float* t1, float* out,
                       int B, int T, bvg_stream_t stream);
/* options: "graph" (0 never, 1 always, 2 [default] from the second forward of a (B, T) shape on: the launch sequence replays
 * as one CUDA graph - the solver calls the estimator 25 times per utterance with one shape), "conv_own_sm" (as bvg_set_option) */
int bvg_s2mel_tail_set_option(bvg_s2mel_tail* h, const char* key, int value);
int bvg_s2mel_tail_last_forward_launches(const bvg_s2mel_tail* h);
/* One Euler step of BASECFM.solve_euler in place on x [B, C, T] fp32 (device):
 *   cfg_rate > 0: dphi is the stacked estimator output [2B, C, T]; d = (1 + cfg_rate) * dphi[:B] - cfg_rate * dphi[B:]
 *   else        : dphi is [B, C, T]; d = dphi
 *   x = x + dt * d;  x[:, :, :prompt_len] = 0            (flow_matching.py:103-112; bit-identical to the fp32 tensor ops) */
int bvg_cfm_euler_step(float* x, const float* dphi, float dt, double cfg_rate, int B, int C, int64_t T, int64_t prompt_len,
                       bvg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BVG_B200_H_ */
