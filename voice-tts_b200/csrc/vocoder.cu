// Whole-generator plan: weight ingestion (reference state-dict keys), workspace
// arena, and the layer sequence of BigVGAN.forward (bigvgan.py:360-386) expressed
// over the channels-last kernels of this library.  Host C++ only - no torch.
//
// Data flow per upsampling stage (bf16 mode; fp32 mode stores everything in fp32):
//   X  (fp32) = ConvTranspose1d(prev)                      residual stream, shared by the 3 AMP blocks
//   per AMP block j, per dilation layer l:
//     A1 (bf16) = Activation1d(cur)                        cur = X (l == 0) or Y
//     M  (bf16) = Conv1d(A1, dil = d_l) + bias
//     A2 (bf16) = Activation1d(M)
//     Y  (fp32) = Conv1d(A2) + bias + cur                  (l < last)
//     XS (fp32) = (Conv1d(A2) + bias + cur) / nk [+ XS]    (l == last; the (xs0+xs1+xs2)/3 of bigvgan.py:369-375
//                                                           folded into the epilogue; the last block of a stage
//                                                           writes the next stage's bf16 ConvTranspose input)
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "conv.cuh"

namespace bvg {

struct ConvW {
  void* ws[3] = {nullptr, nullptr, nullptr};   // fp32 mode: the weight as three bf16 terms w = w0 + w1 + w2, each packed like `w`
  void* w = nullptr;       // packed Wp[k][Cout_r][Cin_p]
  float* bias = nullptr;   // [Cout_r]
  int Cin = 0, Cout = 0, Cin_p = 0, Cout_p = 0, Cout_n = 0, Cout_r = 0;
  int k = 0, k_torch = 0, dil = 1, up = 0;  // up > 0: ConvTranspose1d with stride `up`
  bool has_w = false, has_b = false, want_b = true;
  // Time folding of narrow layers (bf16 mode, see build_fold_twin): a Conv1d over [T, Cp] with Cp <= 64 is the SAME memory as
  // [T / F, F * Cp] and the same arithmetic as a conv with kf taps over F * Cp "phase channels" - the MMA tile (M = 128) then
  // holds F time phases instead of F replicas of the layer.  The twin carries the re-packed weights / bias.
  std::vector<float> w_host, b_host;  // torch-layout fp32 copies, kept only for layers that may fold
  ConvW* fold_twin = nullptr;
  int fold = 1;
};

// 1x1 conv on the speaker embedding (models.py:204-209): weight [C][E] fp32, bias [C]
struct CondW {
  float* w = nullptr;
  float* b = nullptr;
  int C = 0;
  bool has_w = false, has_b = false;
};

struct ActW {
  float* alpha = nullptr;  // [Cp] log scale
  float* beta = nullptr;   // [Cp] log scale
  Taps taps;
  int C = 0, Cp = 0;
  bool has_a = false, has_b = false, has_up = false, has_down = false;
};

}  // namespace bvg

using namespace bvg;

struct bvg_vocoder {
  bvg_config cfg;
  int nst = 0, nk = 0, nd = 0;
  int act_dt = BVG_BF16;          // storage type of conv operands
  std::vector<int> C, Cp;         // C[0] = initial channels, C[i+1] = stage i
  int mel_p = 0;
  int total_up = 1;
  ConvW conv_pre;
  std::vector<ConvW> ups;
  std::vector<ConvW> convs1, convs2;   // [(stage*nk + j)*nd + l]
  std::vector<ActW> acts;              // [(stage*nk + j)*2*nd + a]
  ActW act_post;
  CondW cond_pre;                      // cond_layer (added after conv_pre)
  std::vector<CondW> conds;            // conds.{i} (added after ups[i]); empty unless cond_each_up
  int E = 0;                           // speaker-embedding width (0: unconditioned v2 generator)
  float* post_w = nullptr;             // [7][Cp_last]
  float post_bias = 0.f;
  bool has_post_w = false, has_post_b = false;
  bool finalized = false;
  // options
  int opt_graph = 2;               // CUDA-graph replay of the layer sequence: 0 never, 1 from the first forward of a shape, 2 (default) from the SECOND
                                   // forward of a (B, T0) shape on (repeated shapes - a serving loop, the benchmark - replay; one-off shapes never pay a capture)
  int opt_conv_impl = 0, opt_fast_sin = -1;
  int opt_own_sm = 1;              // persistent tcgen05 conv CTAs take their SM's whole shared-memory carve-out (ConvArgs::own_sm)
  int opt_split_terms = 3;         // fp32 storage + tensor cores (conv_impl = 3): term pairs per convolution (3, 6 or 9)
  int opt_fuse_res = 1;            // conv2 of an AMP unit adds the residual AND applies the next unit's first activation (bf16 mode)
  int opt_fuse_act = 1;            // conv1 of an AMP unit applies the following activation in its epilogue (bf16 mode)
  int fuse_res_min_kc = 4096;      // smallest k * Cin whose conv2 takes the fused residual + activation epilogue
  int opt_fuse_unit = 0;           // 1: whole AMP units of <= 96-channel stages as one kernel (amp_unit.cu; bf16 mode) - measured slower
                                   // than the layer-by-layer path on B200, so off by default
  int opt_fold = 1;                // narrow resblock convolutions as time-folded F * Cp-channel layers where that saves MMAs
  int opt_pdl = 0;                 // programmatic dependent launch of the conv / activation kernels (common.cuh): bit-identical,
                                   // measured no gain (16 x 10 s: 39.8 vs 39.6 ms; 1 x 2 s: 1.44 vs 1.35 ms) - off by default
  int opt_streams = 3;             // AMP blocks of one stage run on up to this many streams (1 = serial); see DESIGN.md 8.5
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // internal streams for AMP blocks 0 .. nk-2
  cudaEvent_t ev_fork = nullptr, ev_blk[3] = {nullptr, nullptr, nullptr};
  int64_t opt_ws_cap_mb = 24 * 1024;
  // workspace
  void* arena = nullptr;
  size_t arena_bytes = 0;
  float* pin_mel = nullptr; size_t pin_mel_bytes = 0;
  void* pin_wav = nullptr;  size_t pin_wav_bytes = 0;
  float* dev_mel = nullptr; size_t dev_mel_bytes = 0;
  void* dev_wav = nullptr;  size_t dev_wav_bytes = 0;
  int last_launches = 0;
  // per-kernel CUDA-event profiling (option "profile"): category -> accumulated work; events resolved on read
  // consecutive launches on one stream share the boundary event (end of launch i = start of launch i+1)
  struct ProfRec { int cat; double work; int i0, i1; int cin, cout, k, dil; long long rows; };
  int opt_profile = 0;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_ev;      // events of the current window, indexed by ProfRec::i0/i1
  int prof_last = -1;                    // index of the last boundary event, -1: none
  cudaStream_t prof_last_stream = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  std::map<std::pair<int, int>, std::pair<cudaGraphExec_t, int>> graphs;  // exec + kernels inside
  std::map<std::pair<int, int>, int> shape_seen;                          // forwards per (B, T0) so far (graph = 2)
};

static int drop_graphs(bvg_vocoder* v);   // defined with bvg_set_option below

namespace bvg {

static int dev_alloc(void** p, size_t bytes) {
  BVG_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  return BVG_OK;
}

static void init_conv(ConvW& c, int Cin, int Cout, int k, int dil, int up, int gran) {
  c.Cin = Cin; c.Cout = Cout; c.k_torch = k; c.dil = dil; c.up = up;
  c.Cin_p = pad_channels(Cin, gran);
  c.Cout_p = pad_channels(Cout, gran);
  if (up > 0) {
    c.k = 3; c.dil = 1;
    c.Cout_n = up * c.Cout_p;
  } else {
    c.k = k;
    c.Cout_n = c.Cout_p;
  }
  c.Cout_r = round_up(c.Cout_n, 128);
}

static int alloc_conv(ConvW& c, int w_dt) {
  int rc = dev_alloc(&c.w, (size_t)c.k * c.Cout_r * c.Cin_p * dtype_size(w_dt));
  if (rc) return rc;
  if (w_dt == BVG_F32)
    for (int i = 0; i < 3; ++i)
      if ((rc = dev_alloc(&c.ws[i], (size_t)c.k * c.Cout_r * c.Cin_p * 2))) return rc;
  rc = dev_alloc((void**)&c.bias, (size_t)c.Cout_r * sizeof(float));
  if (rc) return rc;
  BVG_CUDA(cudaMemset(c.bias, 0, (size_t)c.Cout_r * sizeof(float)));
  return BVG_OK;
}

static int alloc_act(ActW& a, int C, int gran) {
  a.C = C; a.Cp = pad_channels(C, gran);
  int rc = dev_alloc((void**)&a.alpha, a.Cp * sizeof(float));
  if (rc) return rc;
  rc = dev_alloc((void**)&a.beta, a.Cp * sizeof(float));
  if (rc) return rc;
  BVG_CUDA(cudaMemset(a.alpha, 0, a.Cp * sizeof(float)));
  BVG_CUDA(cudaMemset(a.beta, 0, a.Cp * sizeof(float)));
  return BVG_OK;
}

// a1 / m / a2 / y exist once per concurrently running AMP block (nb sets)
struct Buffers {
  float *cb_pre, *cb_up[8];            // per-utterance bias rows of conv_pre / ups[i] (conditioned generator only)
  void *mel, *p0, *nx, *a1[4], *m[4], *a2[4];
  float *x, *y[4], *y2[4], *xs;   // y / y2: the residual stream of a block ping-pongs (fused conv2 reads halo rows of its input)
  void* sp[4][3];                 // fp32 mode: the conv input as three bf16 terms (per concurrently running block)
  float* t[4];                    // fp32 mode: running sum of the term-pair convolutions
  int nb;
};

static int n_block_streams(const bvg_vocoder* v) {
  int n = v->opt_streams < 1 ? 1 : v->opt_streams;
  if (n > v->nk) n = v->nk;
  return n > 4 ? 4 : n;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace layout for a micro-batch of B utterances of T0 frames
static size_t plan_buffers(const bvg_vocoder* v, int B, int T0, Buffers* out) {
  const size_t es = dtype_size(v->act_dt);
  size_t nmax = 0;
  int64_t T = T0;
  for (int i = 0; i < v->nst; ++i) {
    T *= v->cfg.upsample_rates[i];
    const size_t n = (size_t)B * T * v->Cp[i + 1];
    if (n > nmax) nmax = n;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  size_t o_cb_pre = 0, o_cb_up[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (v->E > 0) {
    o_cb_pre = take((size_t)B * v->conv_pre.Cout_r * 4);
    if (!v->conds.empty())
      for (int i = 0; i < v->nst; ++i) o_cb_up[i] = take((size_t)B * v->ups[i].Cout_r * 4);
  }
  const size_t o_mel = take((size_t)B * T0 * v->mel_p * es);
  const size_t o_p0 = take((size_t)B * T0 * v->Cp[0] * es);
  const size_t o_nx = take(nmax * es);
  const size_t o_x = take(nmax * 4);
  const size_t o_xs = take(nmax * 4);
  const int nb = n_block_streams(v);
  size_t o_a1[4], o_m[4], o_a2[4], o_y[4], o_y2[4], o_sp[4][3] = {}, o_t[4] = {};
  const bool split = v->cfg.mode == BVG_MODE_FP32;
  size_t nin = nmax;                                  // largest conv input: stage tensors, conv_pre's mel, ups[0]'s p0
  if ((size_t)B * T0 * v->mel_p > nin) nin = (size_t)B * T0 * v->mel_p;
  if ((size_t)B * T0 * v->Cp[0] > nin) nin = (size_t)B * T0 * v->Cp[0];
  for (int j = 0; j < nb; ++j) {
    o_a1[j] = take(nmax * es);
    o_m[j] = take(nmax * es);
    o_a2[j] = take(nmax * es);
    o_y[j] = take(nmax * 4);
    o_y2[j] = take(nmax * 4);
    if (split) {
      for (int q = 0; q < 3; ++q) o_sp[j][q] = take(nin * 2);
      o_t[j] = take((nmax > (size_t)B * T0 * v->Cp[0] ? nmax : (size_t)B * T0 * v->Cp[0]) * 4);
    }
  }
  if (out) {
    unsigned char* base = (unsigned char*)v->arena;
    out->cb_pre = (float*)(base + o_cb_pre);
    for (int i = 0; i < 8; ++i) out->cb_up[i] = (float*)(base + o_cb_up[i]);
    out->mel = base + o_mel; out->p0 = base + o_p0; out->nx = base + o_nx;
    out->x = (float*)(base + o_x); out->xs = (float*)(base + o_xs);
    out->nb = nb;
    for (int j = 0; j < nb; ++j) {
      out->a1[j] = base + o_a1[j]; out->m[j] = base + o_m[j]; out->a2[j] = base + o_a2[j];
      out->y[j] = (float*)(base + o_y[j]);
      out->y2[j] = (float*)(base + o_y2[j]);
      for (int q = 0; q < 3; ++q) out->sp[j][q] = split ? base + o_sp[j][q] : nullptr;
      out->t[j] = split ? (float*)(base + o_t[j]) : nullptr;
    }
  }
  return off;
}

static int max_microbatch(const bvg_vocoder* v, int B, int T0) {
  const size_t cap = (size_t)v->opt_ws_cap_mb << 20;
  int b = B;
  while (b > 1 && plan_buffers(v, b, T0, nullptr) > cap) b = (b + 1) / 2;
  return b;
}

enum { CAT_CONV_UMMA = 0, CAT_CONV_SIMT = 1, CAT_ACT = 2, CAT_OTHER = 3, CAT_UNIT = 4, CAT_N = 5 };

static cudaEvent_t prof_event(bvg_vocoder* v) {
  cudaEvent_t e = nullptr;
  if (!v->ev_pool.empty()) { e = v->ev_pool.back(); v->ev_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  bvg_vocoder* v; cudaStream_t st; int cat; double work; int i0 = -1;
  int cin = 0, cout = 0, k = 0, dil = 0; long long rows = 0;
  ProfScope(bvg_vocoder* v_, cudaStream_t st_, int cat_, double work_) : v(v_), st(st_), cat(cat_), work(work_) {
    if (!v->opt_profile) return;
    if (v->prof_last >= 0 && v->prof_last_stream == st) {
      i0 = v->prof_last;                 // nothing was enqueued on this stream since the previous launch ended
    } else {
      cudaEvent_t e = prof_event(v); cudaEventRecord(e, st);
      v->prof_ev.push_back(e); i0 = (int)v->prof_ev.size() - 1;
    }
  }
  ~ProfScope() {
    if (i0 < 0) return;
    cudaEvent_t e = prof_event(v); cudaEventRecord(e, st);
    v->prof_ev.push_back(e);
    const int i1 = (int)v->prof_ev.size() - 1;
    v->prof.push_back({cat, work, i0, i1, cin, cout, k, dil, rows});
    v->prof_last = i1; v->prof_last_stream = st;
  }
};
// anything enqueued outside a ProfScope (copies, event waits, graph launches) breaks the chain of shared events
static inline void prof_break(bvg_vocoder* v) { v->prof_last = -1; }

// fp32 storage on the tensor cores (bvg_set_option("conv_impl", 3) in BVG_MODE_FP32; Python precision "bf16x3"): activation
// and weight are each the sum of three bf16 terms (exact to fp32's 24 bits) and the convolution is the sum of the term-pair
// convolutions, (a0+a1+a2)(w0+w1+w2) ~ [a2w2 + a2w1 + a1w2 +] [a2w0 + a0w2 + a1w1 +] a1w0 + a0w1 + a0w0 - "split_terms" 9,
// 6 or 3 (default) of them, smallest first - every pair one bf16 tcgen05 launch accumulating in fp32 (TMEM), chained through
// the fp32 running sum `t`; bias / residual / scale / accumulate ride on the first and last pass.  bf16 x bf16 products are
// exact in fp32, so with 6 or 9 terms the operands carry full fp32 precision - and the result still differs from the fp32
// SIMT kernels by 1.8e-5 on the full generator (95.7 dB; SIMT 2.6e-6 / 112 dB): the tensor cores' fp32 accumulation is not
// round-to-nearest, and a 768-channel k = 11 layer chains 528 accumulating MMAs per output.  That is why this is NOT the
// <= 1e-5 parity mode, and why 3 terms (93.4 dB) are the default: measured on 16 x 10 s, 3 / 6 / 9 terms run at 1 150 / 674 /
// 478 audio-s/s against 147-223 for the SIMT kernels (first / second generation) and 4 000 for plain bf16 (41 dB).
struct SplitPlan { ConvArgs first, mid, last; };
static bool plan_conv_split(const ConvW& c, void* out, int out_dt, const float* res, const float* accum, float scale, int B,
                            int64_t T, const float* bias, int64_t bias_bs, void* const sp[3], float* t, SplitPlan* pl, int own_sm) {
  ConvArgs a;
  a.own_sm = own_sm;
  a.in = sp[0]; a.w = c.ws[0]; a.bias = nullptr; a.out = t; a.res = res; a.accum = nullptr; a.scale = 1.f;
  a.in_dtype = BVG_BF16; a.w_dtype = BVG_BF16; a.out_dtype = BVG_F32;
  a.B = B; a.T = T; a.Cin_p = c.Cin_p; a.Cout_n = c.Cout_n; a.Cout_r = c.Cout_r; a.out_ld = c.Cout_n;
  a.k = c.k; a.dil = c.dil;
  pl->first = a;
  pl->mid = a;
  pl->mid.res = t;
  pl->last = a;
  pl->last.res = t; pl->last.accum = accum; pl->last.out = out; pl->last.out_dtype = out_dt; pl->last.bias = bias;
  pl->last.bias_bs = bias_bs; pl->last.scale = scale;
  return conv_umma_supported(pl->first) && conv_umma_supported(pl->mid) && conv_umma_supported(pl->last);
}
static int run_conv_split(bvg_vocoder* v, const ConvW& c, const float* in, SplitPlan& pl, int B, int64_t T, cudaStream_t st,
                          void* const sp[3]) {
  int rc = split3_bf16(sp[0], sp[1], sp[2], in, (int64_t)B * T * c.Cin_p, st);
  if (rc) return rc;
  // (activation term, weight term), smallest products first: split_terms = 9 runs all, 6 drops the pairs of weight <= 2^-24,
  // 3 keeps a1w0 + a0w1 + a0w0 (weight >= 2^-8)
  static const int order[9][2] = {{2, 2}, {2, 1}, {1, 2}, {2, 0}, {0, 2}, {1, 1}, {1, 0}, {0, 1}, {0, 0}};
  const int q0 = v->opt_split_terms >= 9 ? 0 : (v->opt_split_terms >= 6 ? 3 : 6);   // 9, 6 or 3 term pairs
  for (int q = q0; q < 9; ++q) {
    ConvArgs& p = q == q0 ? pl.first : (q == 8 ? pl.last : pl.mid);
    p.in = sp[order[q][0]];
    p.w = c.ws[order[q][1]];
    if ((rc = conv_umma_launch(p, 0, st))) return rc;
  }
  return BVG_OK;
}



static int run_conv(bvg_vocoder* v, const ConvW& c, const void* in, int in_dt, void* out, int out_dt,
                    const float* res, const float* accum, float scale, int B, int64_t T, cudaStream_t st,
                    const float* bias_rows = nullptr, void* const* sp = nullptr, float* t = nullptr) {
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.out = out; a.res = res; a.accum = accum; a.scale = scale;
  if (bias_rows) { a.bias = bias_rows; a.bias_bs = c.Cout_r; }   // per-utterance rows: layer bias + cond(speaker_embedding)
  a.in_dtype = in_dt; a.w_dtype = v->act_dt; a.out_dtype = out_dt;
  a.B = B; a.T = T; a.Cin_p = c.Cin_p; a.Cout_n = c.Cout_n; a.Cout_r = c.Cout_r; a.out_ld = c.Cout_n;
  a.k = c.k; a.dil = c.dil; a.own_sm = v->opt_own_sm;
  if (c.fold_twin && v->opt_fold && v->cfg.mode == BVG_MODE_BF16 && v->opt_conv_impl != 1 && in_dt == BVG_BF16 && !bias_rows &&
      T % c.fold == 0) {
    // the same tensors viewed as [T / F, F * Cp], the layer as its time-folded twin (ConvW::fold_twin); flops reported below
    // stay the layer's algorithmic ones
    const ConvW& f = *c.fold_twin;
    ConvArgs b = a;
    b.w = f.w; b.bias = f.bias; b.T = T / c.fold;
    b.Cin_p = f.Cin_p; b.Cout_n = f.Cout_n; b.Cout_r = f.Cout_r; b.out_ld = f.Cout_n; b.k = f.k; b.dil = 1;
    if (conv_umma_supported(b)) a = b;
  }
  if (v->cfg.mode == BVG_MODE_FP32 && v->opt_conv_impl == 3 && in_dt == BVG_F32 && sp && t && c.ws[0]) {
    void* const spl[3] = {sp[0], sp[1], sp[2]};
    SplitPlan pl;
    if (plan_conv_split(c, out, out_dt, res, accum, scale, B, T, a.bias, a.bias_bs, spl, t, &pl, v->opt_own_sm)) {
      ProfScope ps(v, st, CAT_CONV_UMMA, 2.0 * c.Cout * c.Cin * c.k_torch * (double)T * B);
      ps.cin = c.Cin; ps.cout = c.up > 0 ? -c.Cout : c.Cout; ps.k = c.k_torch; ps.dil = 300 + c.dil; ps.rows = (long long)B * T;
      return run_conv_split(v, c, (const float*)in, pl, B, T, st, spl);
    }
  }
  bool umma = (v->cfg.mode == BVG_MODE_BF16) && v->opt_conv_impl != 1;
  if (umma && !conv_umma_supported(a)) {
    if (v->opt_conv_impl == 2) BVG_FAIL(BVG_EINVAL, "conv layer not supported by the tcgen05 kernel");
    umma = false;
  }
  // algorithmic flops: 2*Cout*Cin*k*T_out*B (Conv1d) / 2*Cin*Cout*k*T_in*B (ConvTranspose1d), unpadded channels
  ProfScope ps(v, st, umma ? CAT_CONV_UMMA : CAT_CONV_SIMT, 2.0 * c.Cout * c.Cin * c.k_torch * (double)T * B);
  ps.cin = c.Cin; ps.cout = c.up > 0 ? -c.Cout : c.Cout; ps.k = c.k_torch; ps.dil = c.dil; ps.rows = (long long)B * T;
  const int rc_ = umma ? conv_umma_launch(a, 0, st) : conv_simt_launch(a, st);
  return rc_;
}

// c1 followed by a2 of one AMP unit (bigvgan.py:136-138) as ONE launch when the fused kernel takes the layer
static bool can_fuse_conv_act(const bvg_vocoder* v, const ConvW& c, const void* in, void* out, int B, int64_t T) {
  if (!v->opt_fuse_act || v->cfg.mode != BVG_MODE_BF16 || v->opt_conv_impl == 1 || v->opt_fast_sin == 0) return false;
  // measured on B200 (profiles/r01_layer_times_d.txt, fused vs conv + stand-alone activation): below 192 channels the
  // epilogue activation (two warps per scheduler, a quarter of the TMEM lanes idle) only wins when the MMA stream is long
  // (k = 11); opt_fuse_act = 2 forces the fusion everywhere
  if (v->opt_fuse_act == 1 && c.Cout_n < 192 && c.k < 11) return false;
  // a layer whose time-folded twin runs at a quarter of the MMAs (F = 4: 24-channel k = 11, dilation 1) is faster as folded
  // conv + stand-alone activation (measured round 2: 0.21 + 0.15 ms against 0.53 ms fused)
  if (v->opt_fuse_act == 1 && v->opt_fold && c.fold_twin && c.fold >= 4 && T % c.fold == 0) return false;
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.out = out; a.res = nullptr; a.accum = nullptr; a.scale = 1.f;
  a.in_dtype = BVG_BF16; a.w_dtype = BVG_BF16; a.out_dtype = BVG_BF16;
  a.B = B; a.T = T; a.Cin_p = c.Cin_p; a.Cout_n = c.Cout_n; a.Cout_r = c.Cout_r; a.out_ld = c.Cout_n;
  a.k = c.k; a.dil = c.dil; a.own_sm = v->opt_own_sm;
  return conv_act_fused_supported(a);
}
static int run_conv_act(bvg_vocoder* v, const ConvW& c, const ActW& act, const void* in, void* out, int B, int64_t T,
                        cudaStream_t st) {
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.out = out; a.res = nullptr; a.accum = nullptr; a.scale = 1.f;
  a.in_dtype = BVG_BF16; a.w_dtype = BVG_BF16; a.out_dtype = BVG_BF16;
  a.B = B; a.T = T; a.Cin_p = c.Cin_p; a.Cout_n = c.Cout_n; a.Cout_r = c.Cout_r; a.out_ld = c.Cout_n;
  a.k = c.k; a.dil = c.dil; a.own_sm = v->opt_own_sm;
  // accounted as a conv launch: algorithmic conv flops; the fused activation's algorithmic bytes are zero by construction
  ProfScope ps(v, st, CAT_CONV_UMMA, 2.0 * c.Cout * c.Cin * c.k_torch * (double)T * B);
  ps.cin = c.Cin; ps.cout = c.Cout; ps.k = c.k_torch; ps.dil = 100 + c.dil; ps.rows = (long long)B * T;
  const int rc_ = conv_act_fused_launch(a, act.alpha, act.beta, act.taps, st);
  return rc_;
}

// c2 + residual of unit l and a1 of unit l+1 (bigvgan.py:139 `x = xt + x`, then :134 of the next iteration) as ONE launch:
// a_out = bf16(act(conv + bias + res)), y_out = conv + bias + res (fp32)
static bool can_fuse_conv_res_act(const bvg_vocoder* v, const ConvW& c, const void* in, void* a_out, const float* res,
                                  int B, int64_t T) {
  if (!v->opt_fuse_res || v->cfg.mode != BVG_MODE_BF16 || v->opt_conv_impl == 1 || v->opt_fast_sin == 0) return false;
  // measured on B200 (profiles/r01_layer_times_c.txt): the residual rows come straight from HBM/L2 into registers
  // (one body = 6 steps ahead) and with two epilogue warps per scheduler that latency only hides under long MMA streams:
  // 768 ch k >= 7 and 384 ch k = 11 gain 15-20 %, 384 ch k = 3 and everything narrower lose; fuse_res = 2 forces it
  if (v->opt_fuse_res == 1 && c.k * c.Cin_p < v->fuse_res_min_kc) return false;
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.out = a_out; a.res = res; a.accum = nullptr; a.scale = 1.f;
  a.in_dtype = BVG_BF16; a.w_dtype = BVG_BF16; a.out_dtype = BVG_BF16;
  a.B = B; a.T = T; a.Cin_p = c.Cin_p; a.Cout_n = c.Cout_n; a.Cout_r = c.Cout_r; a.out_ld = c.Cout_n;
  a.k = c.k; a.dil = c.dil; a.own_sm = v->opt_own_sm;
  return conv_act_fused_supported(a);
}
static int run_conv_res_act(bvg_vocoder* v, const ConvW& c, const ActW& act, const void* in, void* a_out, const float* res,
                            float* y_out, int B, int64_t T, cudaStream_t st) {
  ConvArgs a;
  a.in = in; a.w = c.w; a.bias = c.bias; a.out = a_out; a.res = res; a.accum = nullptr; a.scale = 1.f;
  a.in_dtype = BVG_BF16; a.w_dtype = BVG_BF16; a.out_dtype = BVG_BF16;
  a.B = B; a.T = T; a.Cin_p = c.Cin_p; a.Cout_n = c.Cout_n; a.Cout_r = c.Cout_r; a.out_ld = c.Cout_n;
  a.k = c.k; a.dil = c.dil; a.own_sm = v->opt_own_sm;
  ProfScope ps(v, st, CAT_CONV_UMMA, 2.0 * c.Cout * c.Cin * c.k_torch * (double)T * B);
  ps.cin = c.Cin; ps.cout = c.Cout; ps.k = c.k_torch; ps.dil = 200 + c.dil; ps.rows = (long long)B * T;
  const int rc_ = conv_act_fused_launch(a, act.alpha, act.beta, act.taps, st, y_out);
  return rc_;
}

static int run_act(bvg_vocoder* v, const ActW& a, const void* in, int in_dt, void* out, int out_dt, int B,
                   int64_t T, cudaStream_t st) {
  const bool fast = v->opt_fast_sin >= 0 ? v->opt_fast_sin != 0 : v->cfg.mode == BVG_MODE_BF16;
  // algorithmic bytes: one read + one write of the unpadded tensor
  ProfScope ps(v, st, CAT_ACT, (double)B * T * a.C * (dtype_size(in_dt) + dtype_size(out_dt)));
  ps.cin = a.C; ps.cout = a.C; ps.k = (int)dtype_size(in_dt); ps.dil = (int)dtype_size(out_dt); ps.rows = (long long)B * T;
  // with exactly 8 pad channels (the 24 -> 32 channel last stage) only the real channels go through the FIRs and the snake;
  // the thread of the last real pair also stores exact zeros into the pad channels, which the zero weight columns of the
  // next conv need (a stale NaN bit pattern times 0 would poison the accumulator).  Saves a quarter of that stage's work.
  const int Cact = (a.Cp == a.C + 8 && !(a.C & 1) && T >= 64) ? a.C : a.Cp;   // (short sequences: all-scalar launch, all channels)
  const int rc_ = act1d_cl_launch(out, in, a.alpha, a.beta, a.taps, B, T, Cact, in_dt, out_dt, fast, st, a.Cp);
  return rc_;
}

// One whole AMP unit (a1 -> c1 -> a2 -> c2 + residual, bigvgan.py:132-141) as ONE launch for the narrow stages
static void unit_args(const bvg_vocoder* v, int stage, const ConvW& c1, const ConvW& c2, const ActW& a1, const ActW& a2,
                      const float* x, void* out, int out_bf16, const float* accum, float scale, int B, int64_t T,
                      AmpUnitArgs* u) {
  u->x = x; u->out = out; u->accum = accum; u->scale = scale; u->out_bf16 = out_bf16;
  u->w1 = c1.w; u->w2 = c2.w; u->bias1 = c1.bias; u->bias2 = c2.bias;
  u->al1 = a1.alpha; u->be1 = a1.beta; u->al2 = a2.alpha; u->be2 = a2.beta;
  u->taps1 = a1.taps; u->taps2 = a2.taps;
  u->B = B; u->T = T; u->C = v->C[stage + 1]; u->Cp = v->Cp[stage + 1]; u->ld = v->Cp[stage + 1];
  u->k = c1.k; u->dil = c1.dil;
}
static bool can_fuse_unit(const bvg_vocoder* v, int stage, const ConvW& c1, const ConvW& c2, const ActW& a1, const ActW& a2,
                          const float* x, void* out, int B, int64_t T) {
  if (!v->opt_fuse_unit || v->cfg.mode != BVG_MODE_BF16 || v->opt_conv_impl == 1 || v->opt_fast_sin == 0) return false;
  if (c1.Cout_r != 128 || c2.Cout_r != 128 || c1.k != c2.k || c2.dil != 1 || c1.Cin_p != c1.Cout_n) return false;
  AmpUnitArgs u;
  unit_args(v, stage, c1, c2, a1, a2, x, out, 0, nullptr, 1.f, B, T, &u);
  return amp_unit_supported(u);
}
static int run_unit(bvg_vocoder* v, int stage, const ConvW& c1, const ConvW& c2, const ActW& a1, const ActW& a2, const float* x,
                    void* out, int out_bf16, const float* accum, float scale, int B, int64_t T, cudaStream_t st) {
  AmpUnitArgs u;
  unit_args(v, stage, c1, c2, a1, a2, x, out, out_bf16, accum, scale, B, T, &u);
  // work: the algorithmic flops of both convolutions (the two activations cost no algorithmic HBM bytes here)
  ProfScope ps(v, st, CAT_UNIT, 2.0 * 2.0 * c1.Cout * c1.Cin * c1.k_torch * (double)T * B);
  ps.cin = c1.Cin; ps.cout = c1.Cout; ps.k = c1.k_torch; ps.dil = 400 + c1.dil; ps.rows = (long long)B * T;
  return amp_unit_launch(u, st);
}

// internal streams / events for the concurrent AMP blocks (created on first use)
static int ensure_streams(bvg_vocoder* v) {
  for (int i = 0; i < 3; ++i) {
    if (!v->aux[i]) BVG_CUDA(cudaStreamCreateWithFlags(&v->aux[i], cudaStreamNonBlocking));
    if (!v->ev_blk[i]) BVG_CUDA(cudaEventCreateWithFlags(&v->ev_blk[i], cudaEventDisableTiming));
  }
  if (!v->ev_fork) BVG_CUDA(cudaEventCreateWithFlags(&v->ev_fork, cudaEventDisableTiming));
  return BVG_OK;
}

// the layer sequence between the mel transpose and conv_post (graph-capturable)
//
// The nk AMP blocks of a stage read the same X and are independent up to the running sum XS
// (bigvgan.py:369-375), so they are issued on separate streams: the FP32-pipe-bound activation
// kernels of one block run on the SMs beside the tensor-pipe-bound persistent conv CTAs of another.
// Block nk-1 (the longest, k = 11) stays on the caller's stream; block j's last conv waits for block
// j-1's (XS accumulates in block order, so the result is bit-identical to the serial schedule).
static int run_body(bvg_vocoder* v, const Buffers& bf, int B, int T0, cudaStream_t st) {
  const int adt = v->act_dt;
  const int nb = bf.nb;
  int rc = BVG_OK;
  if (nb > 1 && (rc = ensure_streams(v))) return rc;
  const bool cond = v->E > 0;
  void* const* sp_main = bf.sp[nb - 1];   // conv_pre / ups run on the caller's stream, like the last AMP block
  float* t_main = bf.t[nb - 1];
  rc = run_conv(v, v->conv_pre, bf.mel, adt, bf.p0, adt, nullptr, nullptr, 1.f, B, T0, st, cond ? bf.cb_pre : nullptr, sp_main,
                t_main);
  if (rc) return rc;
  const void* stage_in = bf.p0;
  int64_t T = T0;
  for (int i = 0; i < v->nst; ++i) {
    // ConvTranspose1d: 3-tap conv over the input rows writing u*Cp phase channels == [B, u*T, Cp]
    rc = run_conv(v, v->ups[i], stage_in, adt, bf.x, BVG_F32, nullptr, nullptr, 1.f, B, T, st,
                  (cond && !v->conds.empty()) ? bf.cb_up[i] : nullptr, sp_main, t_main);
    if (rc) return rc;
    T *= v->cfg.upsample_rates[i];
    const bool last_stage = (i == v->nst - 1);
    if (nb > 1) BVG_CUDA(cudaEventRecord(v->ev_fork, st));
    for (int j = 0; j < v->nk; ++j) {
      // block -> (stream, buffer set): the last block on the caller's stream, the others round-robin on aux streams
      const int slot = nb > 1 ? (j == v->nk - 1 ? nb - 1 : j % (nb - 1)) : 0;
      cudaStream_t sj = (nb > 1 && j != v->nk - 1) ? v->aux[slot] : st;
      if (nb > 1) prof_break(v);
      if (sj != st) BVG_CUDA(cudaStreamWaitEvent(sj, v->ev_fork, 0));
      const float* cur = bf.x;
      bool a1_ready = false;     // a1 of this unit already came out of the previous unit's fused conv2
      for (int l = 0; l < v->nd; ++l) {
        const int ci = (i * v->nk + j) * v->nd + l;
        const int ai = (i * v->nk + j) * 2 * v->nd + 2 * l;
        if (!a1_ready && can_fuse_unit(v, i, v->convs1[ci], v->convs2[ci], v->acts[ai], v->acts[ai + 1], cur, bf.xs, B, T)) {
          if (l < v->nd - 1) {
            float* ynext = (l & 1) ? bf.y2[slot] : bf.y[slot];
            rc = run_unit(v, i, v->convs1[ci], v->convs2[ci], v->acts[ai], v->acts[ai + 1], cur, ynext, 0, nullptr, 1.f, B, T, sj);
            cur = ynext;
          } else {
            const bool to_next = (j == v->nk - 1) && !last_stage;
            if (nb > 1 && j > 0) { prof_break(v); BVG_CUDA(cudaStreamWaitEvent(sj, v->ev_blk[(j - 1) % 3], 0)); }   // XS of block j-1
            rc = run_unit(v, i, v->convs1[ci], v->convs2[ci], v->acts[ai], v->acts[ai + 1], cur, to_next ? bf.nx : (void*)bf.xs,
                          to_next ? 1 : 0, j > 0 ? bf.xs : nullptr, 1.0f / v->nk, B, T, sj);
            if (rc) return rc;
            if (nb > 1 && j < v->nk - 1) BVG_CUDA(cudaEventRecord(v->ev_blk[j % 3], sj));
          }
          if (rc) return rc;
          continue;
        }
        if (!a1_ready) {
          rc = run_act(v, v->acts[ai], cur, BVG_F32, bf.a1[slot], adt, B, T, sj);
          if (rc) return rc;
        }
        a1_ready = false;
        const size_t nel = (size_t)B * T * v->Cp[i + 1];
        if (can_fuse_conv_act(v, v->convs1[ci], bf.a1[slot], bf.a2[slot], B, T)) {
          rc = run_conv_act(v, v->convs1[ci], v->acts[ai + 1], bf.a1[slot], bf.a2[slot], B, T, sj);
          if (rc) return rc;
        } else {
          rc = run_conv(v, v->convs1[ci], bf.a1[slot], adt, bf.m[slot], adt, nullptr, nullptr, 1.f, B, T, sj, nullptr,
                        bf.sp[slot], bf.t[slot]);
          if (rc) return rc;
          rc = run_act(v, v->acts[ai + 1], bf.m[slot], adt, bf.a2[slot], adt, B, T, sj);
          if (rc) return rc;
        }
        if (l < v->nd - 1) {
          float* ynext = (l & 1) ? bf.y2[slot] : bf.y[slot];
          if (can_fuse_conv_res_act(v, v->convs2[ci], bf.a2[slot], bf.a1[slot], cur, B, T)) {
            rc = run_conv_res_act(v, v->convs2[ci], v->acts[ai + 2], bf.a2[slot], bf.a1[slot], cur, ynext, B, T, sj);
            a1_ready = true;
          } else {
            rc = run_conv(v, v->convs2[ci], bf.a2[slot], adt, ynext, BVG_F32, cur, nullptr, 1.f, B, T, sj, nullptr, bf.sp[slot],
                          bf.t[slot]);
          }
          cur = ynext;
        } else {
          const bool to_next = (j == v->nk - 1) && !last_stage;
          if (nb > 1 && j > 0) { prof_break(v); BVG_CUDA(cudaStreamWaitEvent(sj, v->ev_blk[(j - 1) % 3], 0)); }   // XS of block j-1
          rc = run_conv(v, v->convs2[ci], bf.a2[slot], adt, to_next ? bf.nx : (void*)bf.xs, to_next ? adt : BVG_F32, cur,
                        j > 0 ? bf.xs : nullptr, 1.0f / v->nk, B, T, sj, nullptr, bf.sp[slot], bf.t[slot]);
          if (rc) return rc;
          if (nb > 1 && j < v->nk - 1) BVG_CUDA(cudaEventRecord(v->ev_blk[j % 3], sj));
        }
        if (rc) return rc;
      }
    }
    stage_in = bf.nx;
  }
  return run_act(v, v->act_post, bf.xs, BVG_F32, bf.a1[0], adt, B, T, st);
}

static int forward_chunk(bvg_vocoder* v, const float* mel, const float* emb, void* wav, int wav_i16, int B, int T0,
                         cudaStream_t st) {
  Buffers bf;
  plan_buffers(v, B, T0, &bf);
  int rc;
  {
    ProfScope ps(v, st, CAT_OTHER, 0.0);
    rc = v->cfg.input_channels_last ? btc_pad_cast(bf.mel, v->act_dt, mel, (int64_t)B * T0, v->cfg.num_mels, v->mel_p, st)
                                    : bct_to_btc(bf.mel, v->act_dt, mel, B, v->cfg.num_mels, v->mel_p, T0, st);
  }
  if (rc) return rc;
  if (v->E > 0) {
    // per-utterance bias rows = layer bias + cond(speaker_embedding)  (models.py:224 `x + self.cond_layer(e)`, :233-234)
    prof_break(v);
    const ConvW& cp = v->conv_pre;
    BVG_CUDA(cudaMemsetAsync(bf.cb_pre, 0, (size_t)B * cp.Cout_r * 4, st));
    rc = cond_bias_launch(bf.cb_pre, cp.Cout_r, cp.bias, v->cond_pre.w, v->cond_pre.b, emb, B, v->E, cp.Cout, cp.Cout_p, 1, st);
    if (rc) return rc;
    for (size_t i = 0; i < v->conds.size(); ++i) {
      const ConvW& cu = v->ups[i];
      BVG_CUDA(cudaMemsetAsync(bf.cb_up[i], 0, (size_t)B * cu.Cout_r * 4, st));
      rc = cond_bias_launch(bf.cb_up[i], cu.Cout_r, cu.bias, v->conds[i].w, v->conds[i].b, emb, B, v->E, cu.Cout, cu.Cout_p,
                            cu.up, st);
      if (rc) return rc;
    }
  }
  bool use_graph = v->opt_graph == 1 && !v->opt_profile;
  if (v->opt_graph == 2 && !v->opt_profile) {
    if (v->shape_seen.size() > 4096) v->shape_seen.clear();
    use_graph = v->shape_seen[std::make_pair(B, T0)]++ >= 1;
  }
  if (use_graph) {
    auto key = std::make_pair(B, T0);
    auto it = v->graphs.find(key);
    if (it == v->graphs.end()) {
      if (v->graphs.size() >= 32) {              // bound the cache: a graph holds ~200 kernel nodes
        BVG_CUDA(cudaDeviceSynchronize());        // replays of the old graphs may still be in flight on other streams
        for (auto& kv : v->graphs) cudaGraphExecDestroy(kv.second.first);
        v->graphs.clear();
      }
      cudaGraph_t g = nullptr;
      cudaStream_t cs;
      BVG_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      BVG_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      const uint64_t k0 = g_launches.load();
      rc = run_body(v, bf, B, T0, cs);
      const int nkern = (int)(g_launches.load() - k0);
      g_launches.store(k0);  // captured, not launched
      cudaError_t e = cudaStreamEndCapture(cs, &g);
      cudaStreamDestroy(cs);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      BVG_CUDA(e);
      cudaGraphExec_t ge = nullptr;
      BVG_CUDA(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      it = v->graphs.emplace(key, std::make_pair(ge, nkern)).first;
    }
    BVG_CUDA(cudaGraphLaunch(it->second.first, st));
    g_launches.fetch_add((uint64_t)it->second.second);
  } else {
    rc = run_body(v, bf, B, T0, st);
    if (rc) return rc;
  }
  const int64_t Tw = (int64_t)T0 * v->total_up;
  ProfScope ps(v, st, CAT_OTHER, 0.0);
  return conv_post_launch(wav, wav_i16, bf.a1[0], v->act_dt, v->post_w, v->post_bias, B, v->Cp[v->nst], Tw,
                          v->cfg.use_tanh_at_final, st);
}

static int ensure_arena(bvg_vocoder* v, size_t need) {
  if (need <= v->arena_bytes) return BVG_OK;
  // graphs bake arena addresses
  for (auto& kv : v->graphs) cudaGraphExecDestroy(kv.second.first);
  v->graphs.clear();
  if (v->arena) {
    BVG_CUDA(cudaDeviceSynchronize());
    cudaFree(v->arena);
    v->arena = nullptr;
    v->arena_bytes = 0;
  }
  BVG_CUDA(cudaMalloc(&v->arena, need));
  v->arena_bytes = need;
  return BVG_OK;
}

int vocoder_forward(bvg_vocoder* v, const float* mel, const float* emb, void* wav, int wav_i16, int B, int T0,
                    cudaStream_t st) {
  if (!v || !v->finalized) BVG_FAIL(BVG_ESTATE, "vocoder handle is not finalized");
  if (v->E > 0 && !emb) BVG_FAIL(BVG_EINVAL, "speaker-conditioned generator: call bvg_vocoder_fwd_cond with the speaker embedding");
  if (v->E == 0 && emb) BVG_FAIL(BVG_EINVAL, "this generator takes no speaker embedding (cond_dim = 0)");
  if (B < 0 || T0 < 0) BVG_FAIL(BVG_EINVAL, "negative batch or length");
  if (B == 0 || T0 == 0) return BVG_OK;
  if (!mel || !wav) BVG_FAIL(BVG_EINVAL, "null mel/wav pointer");
  if ((int64_t)T0 * v->total_up > 0x3fffffffLL) BVG_FAIL(BVG_EINVAL, "utterance too long");
  BVG_DEVICE(v->cfg.device);
  int rc = ensure_device_ok();
  if (rc) return rc;
  const uint64_t l0 = g_launches.load();
  PdlScope pdl_scope(v->opt_pdl && !v->opt_profile);   // per-launch profiling events want plain stream order
  prof_break(v);   // the caller may have enqueued work on `st` since the last forward
  const int mb = max_microbatch(v, B, T0);
  rc = ensure_arena(v, plan_buffers(v, mb, T0, nullptr));
  if (rc) return rc;
  const int64_t Tw = (int64_t)T0 * v->total_up;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int bc = (B - b0 < mb) ? B - b0 : mb;
    void* wv = wav_i16 ? (void*)((int16_t*)wav + (int64_t)b0 * Tw) : (void*)((float*)wav + (int64_t)b0 * Tw);
    rc = forward_chunk(v, mel + (int64_t)b0 * v->cfg.num_mels * T0, emb ? emb + (int64_t)b0 * v->E : nullptr, wv, wav_i16, bc,
                       T0, st);
    if (rc) return rc;
  }
  v->last_launches = (int)(g_launches.load() - l0);
  return BVG_OK;
}

// ------------------------------------------------------------------ weights ----
static bool parse_int(const char*& s, int* out) {
  if (*s < '0' || *s > '9') return false;
  int v = 0, nd = 0;
  while (*s >= '0' && *s <= '9') {
    if (++nd > 6) return false;   // no layer index has more digits; keeps the arithmetic far from overflow
    v = v * 10 + (*s++ - '0');
  }
  *out = v;
  return true;
}
static bool eat(const char*& s, const char* lit) {
  size_t n = strlen(lit);
  if (strncmp(s, lit, n) != 0) return false;
  s += n;
  return true;
}

// ---- time folding of narrow resblock convolutions -----------------------------------------------------------------
// out[t][co] = sum_j sum_ci W[co][ci][j] x[t + (j - c) d][ci]   with t = F t' + po and input row F (t' + tau) + pi:
//   F tau + pi - po = (j - c) d   =>   Wf[tau][(po, co)][(pi, ci)] = W[co][ci][c + (F tau + pi - po) / d]   (0 where undefined)
// a conv with kf = 2 floor((c d + F - 1) / F) + 1 taps, dilation 1, over F * Cp channels and T / F rows of the SAME tensors
// (channels-last [T, Cp] is [T / F, F Cp]; pad channels stay pad channels; zero fill outside [0, T / F) is the layer's own
// zero padding).  tcgen05 pays per K = 16 step whatever M holds, so the fold pays when kf * F Cp / 16 is well under
// k * F * ceil(Cp / 16): 24 channels k = 11: 40 MMAs per 1 024 samples instead of 88, k = 7: 24 instead of 56.
static int fold_factor(const bvg_vocoder* v, const ConvW& c) {
  if (v->act_dt != BVG_BF16 || c.up > 0 || c.Cin_p != c.Cout_p || c.Cin_p > 64 || c.Cout_n != c.Cout_p) return 1;
  const int F = 128 / c.Cin_p;
  if (F < 2 || (F * c.Cin_p) % 16) return 1;
  const int S = (c.k - 1) / 2 * c.dil, kf = 2 * ((S + F - 1) / F) + 1;
  const int mm_plain = c.k * (int)ceil_div(c.Cin_p, 16) * F, mm_fold = kf * (F * c.Cin_p / 16);
  // F = 4 (24 channels) also folds at EQUAL MMA counts - k = 3 with dilation 1 / 3, k = 7 with dilation 3: the folded tile moves
  // 4x the bytes per TMA operation and per tile (measured 0.144 -> 0.110 and 0.202 -> 0.150 ms per launch); at F = 2 that gains nothing
  return mm_fold * 10 <= mm_plain * (F >= 4 ? 10 : 8) ? F : 1;
}

static void free_fold_twin(ConvW& c) {
  if (!c.fold_twin) return;
  if (c.fold_twin->w) cudaFree(c.fold_twin->w);
  if (c.fold_twin->bias) cudaFree(c.fold_twin->bias);
  delete c.fold_twin;
  c.fold_twin = nullptr;
  c.fold = 1;
}

static int build_fold_twin(bvg_vocoder* v, ConvW& c) {
  free_fold_twin(c);
  const int F = fold_factor(v, c);
  if (F <= 1 || c.w_host.empty()) return BVG_OK;
  const int Cp = c.Cin_p, Cf = F * Cp, cen = (c.k - 1) / 2, S = cen * c.dil;
  const int kf = 2 * ((S + F - 1) / F) + 1, cf = (kf - 1) / 2;
  std::vector<float> wf((size_t)Cf * Cf * kf, 0.f), bf((size_t)Cf, 0.f);
  for (int po = 0; po < F; ++po)
    for (int pi = 0; pi < F; ++pi)
      for (int tau = -cf; tau <= cf; ++tau) {
        const int delta = F * tau + pi - po;
        if (delta % c.dil) continue;
        const int j = cen + delta / c.dil;
        if (j < 0 || j >= c.k) continue;
        for (int co = 0; co < c.Cout; ++co)
          for (int ci = 0; ci < c.Cin; ++ci)
            wf[((size_t)(po * Cp + co) * Cf + (pi * Cp + ci)) * kf + (tau + cf)] = c.w_host[((size_t)co * c.Cin + ci) * c.k + j];
      }
  if (!c.b_host.empty())
    for (int po = 0; po < F; ++po)
      for (int co = 0; co < c.Cout; ++co) bf[(size_t)po * Cp + co] = c.b_host[co];
  ConvW* t = new (std::nothrow) ConvW();
  if (!t) BVG_FAIL(BVG_ENOMEM, "out of host memory");
  t->Cin = t->Cout = Cf; t->Cin_p = t->Cout_p = t->Cout_n = Cf; t->Cout_r = round_up(Cf, 128);
  t->k = t->k_torch = kf; t->dil = 1; t->up = 0;
  int rc = alloc_conv(*t, BVG_BF16);
  float* tmp = nullptr;
  if (!rc && cudaMalloc((void**)&tmp, wf.size() * sizeof(float)) != cudaSuccess) rc = BVG_ENOMEM;
  if (!rc && cudaMemcpy(tmp, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) rc = BVG_ECUDA;
  if (!rc) rc = pack_conv_weight(t->w, BVG_BF16, tmp, Cf, Cf, kf, t->Cout_r, Cf, 0);
  if (!rc && cudaMemcpy(t->bias, bf.data(), bf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) rc = BVG_ECUDA;
  if (cudaDeviceSynchronize() != cudaSuccess && !rc) rc = BVG_ECUDA;
  if (tmp) cudaFree(tmp);
  if (rc) {
    if (t->w) cudaFree(t->w);
    if (t->bias) cudaFree(t->bias);
    delete t;
    if (rc == BVG_ECUDA || rc == BVG_ENOMEM) set_error("building the time-folded twin of a %d-channel k = %d layer failed", c.Cin, c.k);
    return rc;
  }
  t->has_w = t->has_b = true;
  c.fold_twin = t;
  c.fold = F;
  return BVG_OK;
}

static int set_conv_tensor(bvg_vocoder* v, ConvW& c, bool is_weight, const float* d_data, int64_t numel,
                           const char* name) {
  if (is_weight) {
    const int64_t want = (int64_t)c.Cin * c.Cout * c.k_torch;
    if (numel != want) BVG_FAIL(BVG_EINVAL, "%s: expected %lld elements, got %lld", name, (long long)want, (long long)numel);
    int rc = c.up > 0 ? pack_convtr_weight(c.w, v->act_dt, d_data, c.Cin, c.Cout, c.up, c.k_torch, c.Cout_p, c.Cout_r, c.Cin_p, 0)
                      : pack_conv_weight(c.w, v->act_dt, d_data, c.Cout, c.Cin, c.k, c.Cout_r, c.Cin_p, 0);
    if (rc) return rc;
    if (fold_factor(v, c) > 1) {
      c.w_host.resize((size_t)numel);
      BVG_CUDA(cudaMemcpy(c.w_host.data(), d_data, (size_t)numel * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (c.ws[0]) {
      // fp32 mode: w = w0 + w1 + w2 (bf16 terms), each packed for the tcgen05 kernels (run_conv_split)
      float* tmp = nullptr;
      BVG_CUDA(cudaMalloc((void**)&tmp, (size_t)3 * numel * sizeof(float)));
      rc = split3_f32(tmp, tmp + numel, tmp + 2 * numel, d_data, numel, 0);
      for (int q = 0; q < 3 && !rc; ++q)
        rc = c.up > 0 ? pack_convtr_weight(c.ws[q], BVG_BF16, tmp + (size_t)q * numel, c.Cin, c.Cout, c.up, c.k_torch, c.Cout_p,
                                           c.Cout_r, c.Cin_p, 0)
                      : pack_conv_weight(c.ws[q], BVG_BF16, tmp + (size_t)q * numel, c.Cout, c.Cin, c.k, c.Cout_r, c.Cin_p, 0);
      cudaError_t e = cudaDeviceSynchronize();
      cudaFree(tmp);
      if (rc) return rc;
      BVG_CUDA(e);
    }
    c.has_w = true;
  } else {
    if (numel != c.Cout) BVG_FAIL(BVG_EINVAL, "%s: expected %d elements, got %lld", name, c.Cout, (long long)numel);
    if (c.up > 0) {
      for (int r = 0; r < c.up; ++r)
        BVG_CUDA(cudaMemcpy(c.bias + (size_t)r * c.Cout_p, d_data, c.Cout * sizeof(float), cudaMemcpyDeviceToDevice));
    } else {
      BVG_CUDA(cudaMemcpy(c.bias, d_data, c.Cout * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    if (fold_factor(v, c) > 1) {
      c.b_host.resize((size_t)c.Cout);
      BVG_CUDA(cudaMemcpy(c.b_host.data(), d_data, (size_t)c.Cout * sizeof(float), cudaMemcpyDeviceToHost));
    }
    c.has_b = true;
  }
  if (v->finalized && fold_factor(v, c) > 1) {
    // a weight replaced after bvg_finalize: the folded twin follows; captured graphs hold the old twin's addresses
    const int rc = ::drop_graphs(v);
    return rc ? rc : build_fold_twin(v, c);
  }
  return BVG_OK;
}

static int set_act_tensor(bvg_vocoder* v, ActW& a, const char* field, const float* d_data, int64_t numel,
                          const char* name) {
  if (!strcmp(field, "act.alpha") || !strcmp(field, "act.beta")) {
    if (numel != a.C) BVG_FAIL(BVG_EINVAL, "%s: expected %d elements, got %lld", name, a.C, (long long)numel);
    std::vector<float> h(a.C);
    BVG_CUDA(cudaMemcpy(h.data(), d_data, a.C * sizeof(float), cudaMemcpyDeviceToHost));
    if (!v->cfg.snake_logscale)
      for (auto& x : h) x = logf(x);  // the kernels apply exp (as the reference's fused kernel, cuda/activation1d.py:68-72)
    const bool is_alpha = !strcmp(field, "act.alpha");
    if (is_alpha) {
      BVG_CUDA(cudaMemcpy(a.alpha, h.data(), a.C * sizeof(float), cudaMemcpyHostToDevice));
      a.has_a = true;
      if (v->cfg.snake_kind == BVG_SNAKE) {  // Snake: beta == alpha (cuda/activation1d.py:61-62)
        BVG_CUDA(cudaMemcpy(a.beta, h.data(), a.C * sizeof(float), cudaMemcpyHostToDevice));
        a.has_b = true;
      }
    } else {
      if (v->cfg.snake_kind == BVG_SNAKE) BVG_FAIL(BVG_EINVAL, "%s: Snake has no beta", name);
      BVG_CUDA(cudaMemcpy(a.beta, h.data(), a.C * sizeof(float), cudaMemcpyHostToDevice));
      a.has_b = true;
    }
    return BVG_OK;
  }
  const bool up = !strcmp(field, "upsample.filter");
  const bool down = !strcmp(field, "downsample.lowpass.filter");
  if (up || down) {
    if (numel != 12) BVG_FAIL(BVG_EINVAL, "%s: only 12-tap filters are supported (got %lld)", name, (long long)numel);
    float h[12];
    BVG_CUDA(cudaMemcpy(h, d_data, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 12; ++i) {
      if (up) a.taps.up[i] = 2.0f * h[i];  // the x2 zero-stuffing gain of UpSample1d (resample.py:33), exact in fp
      else a.taps.down[i] = h[i];
    }
    (up ? a.has_up : a.has_down) = true;
    return BVG_OK;
  }
  BVG_FAIL(BVG_EINVAL, "unknown tensor name '%s'", name);
}

int vocoder_set_tensor(bvg_vocoder* v, const char* name, const float* data, int64_t numel, int is_device) {
  if (!v || !name || !data || numel <= 0) BVG_FAIL(BVG_EINVAL, "bvg_set_tensor: bad argument");
  BVG_DEVICE(v->cfg.device);
  float* tmp = nullptr;
  const float* d = data;
  if (!is_device) {
    BVG_CUDA(cudaMalloc((void**)&tmp, numel * sizeof(float)));
    cudaError_t e = cudaMemcpy(tmp, data, numel * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(tmp); BVG_CUDA(e); }
    d = tmp;
  }
  int rc = BVG_OK;
  const char* s = name;
  int n = 0, l = 0;
  if (eat(s, "conv_pre.")) {
    rc = !strcmp(s, "weight") ? set_conv_tensor(v, v->conv_pre, true, d, numel, name)
         : !strcmp(s, "bias") ? set_conv_tensor(v, v->conv_pre, false, d, numel, name)
                              : (set_error("unknown tensor name '%s'", name), BVG_EINVAL);
  } else if (eat(s, "ups.")) {
    if (!parse_int(s, &n) || n >= v->nst || !eat(s, ".0.")) { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
    else rc = !strcmp(s, "weight") ? set_conv_tensor(v, v->ups[n], true, d, numel, name)
              : !strcmp(s, "bias") ? set_conv_tensor(v, v->ups[n], false, d, numel, name)
                                   : (set_error("unknown tensor name '%s'", name), BVG_EINVAL);
  } else if (eat(s, "resblocks.")) {
    if (!parse_int(s, &n) || n >= v->nst * v->nk || !eat(s, ".")) { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
    else if (bool second = false; eat(s, "convs1.") || (second = eat(s, "convs2."))) {
      if (!parse_int(s, &l) || l >= v->nd || !eat(s, ".")) { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
      else {
        ConvW& c = (second ? v->convs2 : v->convs1)[n * v->nd + l];
        rc = !strcmp(s, "weight") ? set_conv_tensor(v, c, true, d, numel, name)
             : !strcmp(s, "bias") ? set_conv_tensor(v, c, false, d, numel, name)
                                  : (set_error("unknown tensor name '%s'", name), BVG_EINVAL);
      }
    } else if (eat(s, "activations.")) {
      if (!parse_int(s, &l) || l >= 2 * v->nd || !eat(s, ".")) { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
      else rc = set_act_tensor(v, v->acts[n * 2 * v->nd + l], s, d, numel, name);
    } else { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
  } else if (eat(s, "activation_post.")) {
    rc = set_act_tensor(v, v->act_post, s, d, numel, name);
  } else if (v->E > 0 && (!strncmp(s, "cond_layer.", 11) || !strncmp(s, "conds.", 6))) {
    CondW* cw = nullptr;
    if (eat(s, "cond_layer.")) {
      cw = &v->cond_pre;
    } else {
      eat(s, "conds.");
      if (!parse_int(s, &n) || n >= (int)v->conds.size() || !eat(s, ".")) { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
      else cw = &v->conds[n];
    }
    if (cw) {
      const bool is_w = !strcmp(s, "weight");
      if (!is_w && strcmp(s, "bias")) { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
      else if (numel != (is_w ? (int64_t)cw->C * v->E : (int64_t)cw->C)) {
        set_error("%s: expected %lld elements, got %lld", name, (long long)(is_w ? (int64_t)cw->C * v->E : cw->C), (long long)numel);
        rc = BVG_EINVAL;
      } else {   // Conv1d(E, C, 1) weight [C, E, 1] is already the [C][E] matrix
        cudaError_t e2 = cudaMemcpy(is_w ? cw->w : cw->b, d, numel * sizeof(float), cudaMemcpyDeviceToDevice);
        if (e2 != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(e2)); rc = BVG_ECUDA; }
        else (is_w ? cw->has_w : cw->has_b) = true;
      }
    }
  } else if (eat(s, "conv_post.")) {
    const int Cl = v->C[v->nst], Clp = v->Cp[v->nst];
    if (!strcmp(s, "weight")) {
      if (numel != (int64_t)Cl * 7) { set_error("%s: expected %d elements", name, Cl * 7); rc = BVG_EINVAL; }
      else {
        std::vector<float> h(Cl * 7), pk(7 * Clp, 0.f);
        cudaError_t e = cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) {
          for (int c = 0; c < Cl; ++c)
            for (int j = 0; j < 7; ++j) pk[j * Clp + c] = h[c * 7 + j];  // weight [1, C, 7]
          e = cudaMemcpy(v->post_w, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(e)); rc = BVG_ECUDA; }
        else v->has_post_w = true;
      }
    } else if (!strcmp(s, "bias")) {
      if (!v->cfg.use_bias_at_final || numel != 1) { set_error("%s: unexpected conv_post.bias", name); rc = BVG_EINVAL; }
      else {
        cudaError_t e = cudaMemcpy(&v->post_bias, d, sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(e)); rc = BVG_ECUDA; }
        else v->has_post_b = true;
      }
    } else { set_error("unknown tensor name '%s'", name); rc = BVG_EINVAL; }
  } else {
    set_error("unknown tensor name '%s'", name);
    rc = BVG_EINVAL;
  }
  cudaError_t e = cudaDeviceSynchronize();  // packing kernels read `data`; the caller may free it on return
  if (tmp) cudaFree(tmp);
  if (rc == BVG_OK && e != cudaSuccess) { set_error("weight packing failed: %s", cudaGetErrorString(e)); rc = BVG_ECUDA; }
  return rc;
}

int vocoder_create(const bvg_config* cfg, bvg_vocoder** out) {
  if (!cfg || !out) BVG_FAIL(BVG_EINVAL, "bvg_create: null argument");
  *out = nullptr;
  if (cfg->num_upsamples < 1 || cfg->num_upsamples > 8 || cfg->num_kernels < 1 || cfg->num_kernels > 4 ||
      cfg->num_dilations < 1 || cfg->num_dilations > 4 || cfg->num_mels < 1 || cfg->upsample_initial_channel < 2)
    BVG_FAIL(BVG_EINVAL, "bvg_create: configuration out of range");
  if (cfg->mode != BVG_MODE_FP32 && cfg->mode != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_create: unknown mode %d", cfg->mode);
  if (cfg->cond_dim < 0 || cfg->cond_dim > 65536) BVG_FAIL(BVG_EINVAL, "bvg_create: bad cond_dim %d", cfg->cond_dim);
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    const int u = cfg->upsample_rates[i];
    if (!convtr_shape_ok(cfg->upsample_kernel_sizes[i], u))
      BVG_FAIL(BVG_EINVAL, "bvg_create: upsample stage %d needs k - u even and 0 <= (k - u)/2 <= u (got u=%d k=%d)", i, u,
               cfg->upsample_kernel_sizes[i]);
    if ((cfg->upsample_initial_channel >> (i + 1)) < 1) BVG_FAIL(BVG_EINVAL, "bvg_create: too many stages for the channel count");
  }
  for (int j = 0; j < cfg->num_kernels; ++j) {
    if (cfg->resblock_kernel_sizes[j] % 2 != 1) BVG_FAIL(BVG_EINVAL, "bvg_create: resblock kernels must be odd");
    for (int l = 0; l < cfg->num_dilations; ++l)
      if (cfg->resblock_dilations[j][l] < 1) BVG_FAIL(BVG_EINVAL, "bvg_create: bad dilation");
  }
  BVG_DEVICE(cfg->device);
  int rc = ensure_device_ok();
  if (rc) return rc;

  bvg_vocoder* v = new (std::nothrow) bvg_vocoder();
  if (!v) BVG_FAIL(BVG_ENOMEM, "out of host memory");
  v->cfg = *cfg;
  v->nst = cfg->num_upsamples; v->nk = cfg->num_kernels; v->nd = cfg->num_dilations;
  v->act_dt = cfg->mode == BVG_MODE_BF16 ? BVG_BF16 : BVG_F32;
  // channel padding granularity (conv.cuh: pad_channels).  8 is legal on the tcgen05 path, but 48-byte rows (24 channels)
  // halve the TMA load rate of the last stage (measured: its convs 1.7-2.3x slower), so rows stay multiples of 32 bytes
  const int gran = 16;
  v->mel_p = pad_channels(cfg->num_mels, gran);
  v->C.resize(v->nst + 1); v->Cp.resize(v->nst + 1);
  v->C[0] = cfg->upsample_initial_channel;
  for (int i = 0; i < v->nst; ++i) { v->C[i + 1] = cfg->upsample_initial_channel >> (i + 1); v->total_up *= cfg->upsample_rates[i]; }
  for (int i = 0; i <= v->nst; ++i) v->Cp[i] = pad_channels(v->C[i], gran);
  for (int i = 0; i < 12; ++i) { v->act_post.taps.up[i] = 0.f; v->act_post.taps.down[i] = 0.f; }

#define TRY(x) do { rc = (x); if (rc) { bvg_destroy(v); return rc; } } while (0)
  init_conv(v->conv_pre, cfg->num_mels, v->C[0], 7, 1, 0, gran);
  TRY(alloc_conv(v->conv_pre, v->act_dt));
  v->ups.resize(v->nst);
  for (int i = 0; i < v->nst; ++i) {
    init_conv(v->ups[i], v->C[i], v->C[i + 1], cfg->upsample_kernel_sizes[i], 1, cfg->upsample_rates[i], gran);
    TRY(alloc_conv(v->ups[i], v->act_dt));
  }
  v->convs1.resize(v->nst * v->nk * v->nd); v->convs2.resize(v->nst * v->nk * v->nd);
  v->acts.resize(v->nst * v->nk * 2 * v->nd);
  for (int i = 0; i < v->nst; ++i)
    for (int j = 0; j < v->nk; ++j)
      for (int l = 0; l < v->nd; ++l) {
        const int ci = (i * v->nk + j) * v->nd + l;
        init_conv(v->convs1[ci], v->C[i + 1], v->C[i + 1], cfg->resblock_kernel_sizes[j], cfg->resblock_dilations[j][l], 0, gran);
        init_conv(v->convs2[ci], v->C[i + 1], v->C[i + 1], cfg->resblock_kernel_sizes[j], 1, 0, gran);
        TRY(alloc_conv(v->convs1[ci], v->act_dt));
        TRY(alloc_conv(v->convs2[ci], v->act_dt));
        for (int a = 0; a < 2; ++a) TRY(alloc_act(v->acts[(i * v->nk + j) * 2 * v->nd + 2 * l + a], v->C[i + 1], gran));
      }
  TRY(alloc_act(v->act_post, v->C[v->nst], gran));
  TRY(dev_alloc((void**)&v->post_w, 7 * v->Cp[v->nst] * sizeof(float)));
  v->E = cfg->cond_dim;
  if (v->E > 0) {
    auto alloc_cond = [&](CondW& c, int C) -> int {
      c.C = C;
      int r = dev_alloc((void**)&c.w, (size_t)C * v->E * sizeof(float));
      return r ? r : dev_alloc((void**)&c.b, (size_t)C * sizeof(float));
    };
    TRY(alloc_cond(v->cond_pre, v->C[0]));
    if (cfg->cond_each_up) {
      v->conds.resize(v->nst);
      for (int i = 0; i < v->nst; ++i) TRY(alloc_cond(v->conds[i], v->C[i + 1]));
    }
  }
#undef TRY
  *out = v;
  return BVG_OK;
}

int vocoder_finalize(bvg_vocoder* v) {
  if (!v) BVG_FAIL(BVG_EINVAL, "null handle");
  auto chk_conv = [&](const ConvW& c, const char* what, int idx, bool need_bias) -> int {
    if (!c.has_w) BVG_FAIL(BVG_ESTATE, "missing weight: %s[%d].weight", what, idx);
    if (need_bias && !c.has_b) BVG_FAIL(BVG_ESTATE, "missing weight: %s[%d].bias", what, idx);
    return BVG_OK;
  };
  auto chk_act = [&](const ActW& a, const char* what, int idx) -> int {
    if (!a.has_a || !a.has_b) BVG_FAIL(BVG_ESTATE, "missing snake parameters: %s[%d]", what, idx);
    if (!a.has_up || !a.has_down) BVG_FAIL(BVG_ESTATE, "missing filter taps: %s[%d]", what, idx);
    return BVG_OK;
  };
  int rc;
  if ((rc = chk_conv(v->conv_pre, "conv_pre", 0, true))) return rc;
  for (int i = 0; i < v->nst; ++i) if ((rc = chk_conv(v->ups[i], "ups", i, true))) return rc;
  for (size_t i = 0; i < v->convs1.size(); ++i) {
    if ((rc = chk_conv(v->convs1[i], "convs1", (int)i, true))) return rc;
    if ((rc = chk_conv(v->convs2[i], "convs2", (int)i, true))) return rc;
  }
  for (size_t i = 0; i < v->acts.size(); ++i) if ((rc = chk_act(v->acts[i], "activations", (int)i))) return rc;
  if ((rc = chk_act(v->act_post, "activation_post", 0))) return rc;
  if (v->E > 0) {
    if (!v->cond_pre.has_w || !v->cond_pre.has_b) BVG_FAIL(BVG_ESTATE, "missing weight: cond_layer.{weight,bias}");
    for (size_t i = 0; i < v->conds.size(); ++i)
      if (!v->conds[i].has_w || !v->conds[i].has_b) BVG_FAIL(BVG_ESTATE, "missing weight: conds.%d.{weight,bias}", (int)i);
  }
  if (!v->has_post_w) BVG_FAIL(BVG_ESTATE, "missing weight: conv_post.weight");
  if (v->cfg.use_bias_at_final && !v->has_post_b) BVG_FAIL(BVG_ESTATE, "missing weight: conv_post.bias");
  for (auto* group : {&v->convs1, &v->convs2})
    for (auto& c : *group) {
      const int rc = build_fold_twin(v, c);
      if (rc) return rc;
    }
  v->finalized = true;
  return BVG_OK;
}

int64_t vocoder_workspace_bytes(const bvg_vocoder* v, int B, int T0) {
  if (!v || B <= 0 || T0 <= 0) return 0;
  return (int64_t)plan_buffers(v, max_microbatch(v, B, T0), T0, nullptr);
}

int vocoder_forward_host(bvg_vocoder* v, const float* mel_host, void* wav_host, int wav_dtype, int B, int T0,
                         cudaStream_t st) {
  if (!v || !v->finalized) BVG_FAIL(BVG_ESTATE, "vocoder handle is not finalized");
  if (B == 0 || T0 == 0) return BVG_OK;
  if (B < 0 || T0 < 0 || !mel_host || !wav_host) BVG_FAIL(BVG_EINVAL, "bad argument");
  if (wav_dtype != 0 && wav_dtype != 1) BVG_FAIL(BVG_EDTYPE, "wav_dtype must be 0 (fp32) or 1 (int16)");
  if (v->E > 0) BVG_FAIL(BVG_EINVAL, "bvg_vocoder_fwd_host: speaker-conditioned generators go through bvg_vocoder_fwd_cond");
  BVG_DEVICE(v->cfg.device);
  const size_t mel_bytes = (size_t)B * v->cfg.num_mels * T0 * sizeof(float);
  const int64_t nw = (int64_t)B * T0 * v->total_up;
  const size_t wav_bytes = (size_t)nw * (wav_dtype ? 2 : 4);
  if (mel_bytes > v->pin_mel_bytes) {
    if (v->pin_mel) cudaFreeHost(v->pin_mel);
    if (v->dev_mel) cudaFree(v->dev_mel);
    v->pin_mel = nullptr; v->dev_mel = nullptr; v->pin_mel_bytes = 0;
    BVG_CUDA(cudaMallocHost((void**)&v->pin_mel, mel_bytes));
    BVG_CUDA(cudaMalloc((void**)&v->dev_mel, mel_bytes));
    v->pin_mel_bytes = mel_bytes;
  }
  if (wav_bytes > v->pin_wav_bytes) {
    if (v->pin_wav) cudaFreeHost(v->pin_wav);
    if (v->dev_wav) cudaFree(v->dev_wav);
    v->pin_wav = nullptr; v->dev_wav = nullptr; v->pin_wav_bytes = 0;
    BVG_CUDA(cudaMallocHost(&v->pin_wav, wav_bytes));
    BVG_CUDA(cudaMalloc(&v->dev_wav, wav_bytes));
    v->pin_wav_bytes = wav_bytes;
  }
  // page-locked caller buffers (cudaHostAlloc / cudaHostRegister, torch pin_memory) are DMA'd directly; pageable ones
  // go through the handle's pinned staging buffers
  auto is_pinned = [](const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  const bool mel_direct = is_pinned(mel_host), wav_direct = is_pinned(wav_host);
  if (!mel_direct) memcpy(v->pin_mel, mel_host, mel_bytes);
  BVG_CUDA(cudaMemcpyAsync(v->dev_mel, mel_direct ? mel_host : v->pin_mel, mel_bytes, cudaMemcpyHostToDevice, st));
  int rc = vocoder_forward(v, v->dev_mel, nullptr, v->dev_wav, wav_dtype, B, T0, st);
  if (rc) return rc;
  BVG_CUDA(cudaMemcpyAsync(wav_direct ? wav_host : v->pin_wav, v->dev_wav, wav_bytes, cudaMemcpyDeviceToHost, st));
  BVG_CUDA(cudaStreamSynchronize(st));
  if (!wav_direct) memcpy(wav_host, v->pin_wav, wav_bytes);
  return BVG_OK;
}

}  // namespace bvg

extern "C" void bvg_destroy(bvg_vocoder* v) {
  if (!v) return;
  DeviceGuard dg(v->cfg.device);
  cudaDeviceSynchronize();
  auto free_conv = [](ConvW& c) {
    free_fold_twin(c);
    if (c.w) cudaFree(c.w);
    if (c.bias) cudaFree(c.bias);
    for (int q = 0; q < 3; ++q) if (c.ws[q]) cudaFree(c.ws[q]);
  };
  auto free_act = [](ActW& a) { if (a.alpha) cudaFree(a.alpha); if (a.beta) cudaFree(a.beta); };
  free_conv(v->conv_pre);
  for (auto& c : v->ups) free_conv(c);
  for (auto& c : v->convs1) free_conv(c);
  for (auto& c : v->convs2) free_conv(c);
  for (auto& a : v->acts) free_act(a);
  free_act(v->act_post);
  if (v->post_w) cudaFree(v->post_w);
  auto free_cond = [](CondW& c) { if (c.w) cudaFree(c.w); if (c.b) cudaFree(c.b); };
  free_cond(v->cond_pre);
  for (auto& c : v->conds) free_cond(c);
  for (auto& e : v->prof_ev) cudaEventDestroy(e);
  for (auto& e : v->ev_pool) cudaEventDestroy(e);
  for (auto& kv : v->graphs) cudaGraphExecDestroy(kv.second.first);
  if (v->arena) cudaFree(v->arena);
  for (int i = 0; i < 3; ++i) { if (v->aux[i]) cudaStreamDestroy(v->aux[i]); if (v->ev_blk[i]) cudaEventDestroy(v->ev_blk[i]); }
  if (v->ev_fork) cudaEventDestroy(v->ev_fork);
  if (v->pin_mel) cudaFreeHost(v->pin_mel);
  if (v->pin_wav) cudaFreeHost(v->pin_wav);
  if (v->dev_mel) cudaFree(v->dev_mel);
  if (v->dev_wav) cudaFree(v->dev_wav);
  delete v;
}

// captured graphs bake the kernel sequence, its launch parameters and the arena addresses: every option that changes what a
// forward enqueues drops them
static int drop_graphs(bvg_vocoder* v) {
  if (v->graphs.empty()) return BVG_OK;
  BVG_DEVICE(v->cfg.device);
  BVG_CUDA(cudaDeviceSynchronize());
  for (auto& kv : v->graphs) cudaGraphExecDestroy(kv.second.first);
  v->graphs.clear();
  return BVG_OK;
}

extern "C" int bvg_set_option(bvg_vocoder* v, const char* key, int value) {
  if (!v || !key) BVG_FAIL(BVG_EINVAL, "bvg_set_option: null argument");
  struct Opt { const char* name; int* field; bool plan; };
  int ws_mb = (int)v->opt_ws_cap_mb;
  const Opt opts[] = {
      {"graph", &v->opt_graph, false},          {"profile", &v->opt_profile, false},
      {"conv_impl", &v->opt_conv_impl, true},   {"split_terms", &v->opt_split_terms, true},
      {"fast_sin", &v->opt_fast_sin, true},     {"workspace_mb", &ws_mb, true},
      {"fuse_res", &v->opt_fuse_res, true},     {"fuse_res_min_kc", &v->fuse_res_min_kc, true},
      {"fuse_act", &v->opt_fuse_act, true},     {"fuse_unit", &v->opt_fuse_unit, true},
      {"streams", &v->opt_streams, true},       {"conv_own_sm", &v->opt_own_sm, true},
      {"pdl", &v->opt_pdl, true},               {"fold", &v->opt_fold, true},
  };
  for (const Opt& o : opts) {
    if (strcmp(key, o.name)) continue;
    if (*o.field == value) return BVG_OK;
    if (o.plan) {
      const int rc = drop_graphs(v);
      if (rc) return rc;
    }
    *o.field = value;
    v->opt_ws_cap_mb = ws_mb;
    return BVG_OK;
  }
  BVG_FAIL(BVG_EINVAL, "bvg_set_option: unknown option '%s'", key);
}

// Debug: one line per recorded launch (category, shape, ms, achieved rate) to `path`; does not clear.
extern "C" int bvg_profile_dump(bvg_vocoder* v, const char* path) {
  if (!v || !path) BVG_FAIL(BVG_EINVAL, "bvg_profile_dump: bad argument");
  BVG_DEVICE(v->cfg.device);
  BVG_CUDA(cudaDeviceSynchronize());
  FILE* f = fopen(path, "w");
  if (!f) BVG_FAIL(BVG_EINVAL, "cannot open %s", path);
  fprintf(f, "cat,cin,cout,k,dil,rows,ms,work,rate\n");
  for (auto& r : v->prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, v->prof_ev[r.i0], v->prof_ev[r.i1]);
    fprintf(f, "%d,%d,%d,%d,%d,%lld,%.4f,%.4g,%.4g\n", r.cat, r.cin, r.cout, r.k, r.dil, r.rows, t, r.work,
            t > 0 ? r.work / (t * 1e-3) : 0.0);
  }
  fclose(f);
  return BVG_OK;
}

// Sums the CUDA-event durations recorded since the last read for one category
// (0 tcgen05 conv, 1 SIMT conv, 2 fused activation, 3 other, 4 whole AMP unit) and clears them when
// category 3 is read (read it last).  Synchronises the device.
extern "C" int bvg_profile_read(bvg_vocoder* v, int category, double* ms, double* work, int* launches) {
  if (!v || category < 0 || category >= CAT_N || !ms || !work || !launches) BVG_FAIL(BVG_EINVAL, "bvg_profile_read: bad argument");
  BVG_DEVICE(v->cfg.device);
  BVG_CUDA(cudaDeviceSynchronize());
  *ms = 0; *work = 0; *launches = 0;
  for (auto& r : v->prof) {
    if (r.cat != category) continue;
    float t = 0.f;
    BVG_CUDA(cudaEventElapsedTime(&t, v->prof_ev[r.i0], v->prof_ev[r.i1]));
    *ms += t; *work += r.work; *launches += 1;
  }
  if (category == CAT_OTHER) {
    for (auto& e : v->prof_ev) v->ev_pool.push_back(e);
    v->prof_ev.clear();
    v->prof.clear();
    v->prof_last = -1;
  }
  return BVG_OK;
}

extern "C" int bvg_last_forward_launches(const bvg_vocoder* v) { return v ? v->last_launches : 0; }
