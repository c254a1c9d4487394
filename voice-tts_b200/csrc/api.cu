// extern "C" entry points of libbvg_b200.so (see include/bvg_b200.h).
#include <cstring>
#include <mutex>

#include "conv.cuh"

struct bvg_vocoder;

namespace bvg {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local int g_pdl = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_device_ok() {
  static std::mutex mu;
  static int ok_dev[64];  // 0 unknown, 1 ok, -1 bad
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    BVG_FAIL(BVG_ENODEV, "no CUDA device available: %s (libbvg_b200 has no CPU fallback)", cudaGetErrorString(e));
  }
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 64 && ok_dev[dev] == 1) return BVG_OK;
  int major = 0, minor = 0;
  BVG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  BVG_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    BVG_FAIL(BVG_ENODEV, "device %d is sm_%d%d; libbvg_b200 is built for sm_100a (B200) only", dev, major, minor);
  if (dev < 64) ok_dev[dev] = 1;
  return BVG_OK;
}

int vocoder_create(const bvg_config* cfg, bvg_vocoder** out);
int vocoder_set_tensor(bvg_vocoder* v, const char* name, const float* data, int64_t numel, int is_device);
int vocoder_finalize(bvg_vocoder* v);
int64_t vocoder_workspace_bytes(const bvg_vocoder* v, int B, int T0);
int vocoder_forward(bvg_vocoder* v, const float* mel, const float* emb, void* wav, int wav_i16, int B, int T0, cudaStream_t st);
int vocoder_forward_host(bvg_vocoder* v, const float* mel_host, void* wav_host, int wav_dtype, int B, int T0,
                         cudaStream_t st);

static void host_taps(Taps* t, const float* up_host, const float* down_host) {
  for (int i = 0; i < 12; ++i) {
    t->up[i] = 2.0f * up_host[i];  // x2 zero-stuffing gain of UpSample1d (resample.py:33); exact in fp
    t->down[i] = down_host[i];
  }
}

static bool on_device(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace bvg

using namespace bvg;

extern "C" {

int bvg_abi_version(void) { return BVG_ABI_VERSION; }
const char* bvg_last_error(void) { return g_err; }
uint64_t bvg_launch_count(void) { return g_launches.load(); }

int bvg_act1d_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log, const float* up_taps,
                  const float* down_taps, int B, int C, int64_t T, int dtype, int flags, bvg_stream_t stream) {
  if (B < 0 || C < 0 || T < 0) BVG_FAIL(BVG_EINVAL, "bvg_act1d_fwd: negative dimension");
  if (dtype != BVG_F32 && dtype != BVG_BF16 && dtype != BVG_F16) BVG_FAIL(BVG_EDTYPE, "bvg_act1d_fwd: unsupported dtype %d", dtype);
  if (B == 0 || C == 0 || T == 0) return BVG_OK;  // reference: seq_len == 0 -> no launch
  if (!dst || !src || !alpha_log || !beta_log || !up_taps || !down_taps) BVG_FAIL(BVG_EINVAL, "bvg_act1d_fwd: null pointer");
  if (dst == src) BVG_FAIL(BVG_EINVAL, "bvg_act1d_fwd: dst must not alias src");
  const size_t es = dtype == BVG_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) % es)
    BVG_FAIL(BVG_EALIGN, "bvg_act1d_fwd: dst/src must be aligned to the element size");
  int rc = ensure_device_ok();
  if (rc) return rc;
  if (!on_device(src) || !on_device(dst)) BVG_FAIL(BVG_EINVAL, "bvg_act1d_fwd: src/dst must be device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  Taps taps;
  host_taps(&taps, up_taps, down_taps);
  return act1d_bct_launch(dst, src, alpha_log, beta_log, taps, B, C, T, dtype, (flags & BVG_ACT_FAST_SIN) != 0, st);
}

int bvg_act1d_cl_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log, const float* up_taps,
                     const float* down_taps, int B, int64_t T, int C, int in_dtype, int out_dtype, int flags,
                     bvg_stream_t stream) {
  if (B < 0 || C < 0 || T < 0) BVG_FAIL(BVG_EINVAL, "bvg_act1d_cl_fwd: negative dimension");
  if ((in_dtype != BVG_F32 && in_dtype != BVG_BF16) || (out_dtype != BVG_F32 && out_dtype != BVG_BF16))
    BVG_FAIL(BVG_EDTYPE, "bvg_act1d_cl_fwd: unsupported dtype");
  if (B == 0 || C == 0 || T == 0) return BVG_OK;
  if (!dst || !src || !alpha_log || !beta_log || !up_taps || !down_taps) BVG_FAIL(BVG_EINVAL, "bvg_act1d_cl_fwd: null pointer");
  if (dst == src) BVG_FAIL(BVG_EINVAL, "bvg_act1d_cl_fwd: dst must not alias src");
  int rc = ensure_device_ok();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Taps taps;
  host_taps(&taps, up_taps, down_taps);
  return act1d_cl_launch(dst, src, alpha_log, beta_log, taps, B, T, C, in_dtype, out_dtype,
                         (flags & BVG_ACT_FAST_SIN) != 0, st);
}

// Stand-alone dense layers on the reference layout: transposes + weight packing
// around the channels-last kernels (test/single-layer use; the vocoder handle
// keeps everything packed and channels-last).
static int dense_layer(float* dst, const float* src, const float* weight, const float* bias, int B, int Cin, int Cout,
                       int64_t T, int k, int dil, int up, int mode, cudaStream_t st, const float* res = nullptr,
                       const float* accum = nullptr, float scale = 1.f, int out_bf16 = 0,
                       const float* act_alpha = nullptr, const float* act_beta = nullptr, const Taps* act_taps = nullptr,
                       float* dst_y = nullptr) {
  if (B < 0 || Cin <= 0 || Cout <= 0 || T < 0 || k <= 0) BVG_FAIL(BVG_EINVAL, "dense layer: bad dimension");
  const int variant = (mode >> 8) & 0xff;  // debug variants ride in the upper bits of `mode`
  mode &= 0xff;
  if (mode != BVG_MODE_FP32 && mode != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "dense layer: unknown mode %d", mode);
  if (B == 0 || T == 0) return BVG_OK;
  if (!dst || !src || !weight) BVG_FAIL(BVG_EINVAL, "dense layer: null pointer");
  if (up > 0 && !convtr_shape_ok(k, up)) BVG_FAIL(BVG_EINVAL, "convtr1d: needs k - stride even and 0 <= (k - stride)/2 <= stride (padding (k - stride)/2, T_out = stride*T)");
  if (up == 0 && (k % 2 != 1 || dil < 1)) BVG_FAIL(BVG_EINVAL, "conv1d: needs odd k and dilation >= 1");
  int rc = ensure_device_ok();
  if (rc) return rc;
  const int dt = mode == BVG_MODE_BF16 ? BVG_BF16 : BVG_F32;
  const size_t es = dtype_size(dt);
  const int gran = 16;   // see vocoder_create
  const int Cin_p = pad_channels(Cin, gran), Cout_p = pad_channels(Cout, gran);
  const int kk = up > 0 ? 3 : k;
  const int Cout_n = up > 0 ? up * Cout_p : Cout_p;
  const int Cout_r = round_up(Cout_n, 128);
  const int64_t Tout = up > 0 ? T * up : T;
  void *xin = nullptr, *wp = nullptr;
  float *bp = nullptr, *yout = nullptr, *resp = nullptr, *accp = nullptr;
  if ((res || accum || out_bf16 || act_taps) && up > 0) BVG_FAIL(BVG_EINVAL, "fused residual / activation forms are defined for conv1d only");
  if (act_taps && (accum || scale != 1.f)) BVG_FAIL(BVG_EINVAL, "fused activation form takes no accumulate operand / scale");
  if (act_taps && ((res != nullptr) != (dst_y != nullptr))) BVG_FAIL(BVG_EINVAL, "fused activation form: residual and y output go together");
  const size_t b_in = (size_t)B * T * Cin_p * es, b_w = (size_t)kk * Cout_r * Cin_p * es, b_b = (size_t)Cout_r * 4,
               b_out = (size_t)B * T * Cout_n * 4;
  unsigned char* blk = nullptr;
  const size_t a = 1024;
  auto up_a = [&](size_t v) { return (v + a - 1) / a * a; };
  BVG_CUDA(cudaMallocAsync((void**)&blk, up_a(b_in) + up_a(b_w) + 3 * up_a(b_b) + 3 * up_a(b_out), st));
  xin = blk; wp = blk + up_a(b_in); bp = (float*)(blk + up_a(b_in) + up_a(b_w));
  yout = (float*)(blk + up_a(b_in) + up_a(b_w) + up_a(b_b));
  resp = (float*)((unsigned char*)yout + up_a(b_out));
  accp = (float*)((unsigned char*)resp + up_a(b_out));
  float* alp = (float*)((unsigned char*)accp + up_a(b_out));   // zero-padded alpha / beta of the fused activation
  float* bep = (float*)((unsigned char*)alp + up_a(b_b));
  do {
    if (act_taps) {
      cudaError_t ea = cudaMemsetAsync(alp, 0, 2 * up_a(b_b), st);
      if (ea == cudaSuccess) ea = cudaMemcpyAsync(alp, act_alpha, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (ea == cudaSuccess) ea = cudaMemcpyAsync(bep, act_beta, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (ea != cudaSuccess) { set_error("activation parameters: %s", cudaGetErrorString(ea)); rc = BVG_ECUDA; break; }
    }
    if ((rc = bct_to_btc(xin, dt, src, B, Cin, Cin_p, T, st))) break;
    if (res && (rc = bct_to_btc(resp, BVG_F32, res, B, Cout, Cout_p, T, st))) break;
    if (accum && (rc = bct_to_btc(accp, BVG_F32, accum, B, Cout, Cout_p, T, st))) break;
    cudaError_t e = cudaMemsetAsync(bp, 0, b_b, st);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); rc = BVG_ECUDA; break; }
    if (up > 0) {
      if ((rc = pack_convtr_weight(wp, dt, weight, Cin, Cout, up, k, Cout_p, Cout_r, Cin_p, st))) break;
      if (bias)
        for (int r = 0; r < up && e == cudaSuccess; ++r)
          e = cudaMemcpyAsync(bp + (size_t)r * Cout_p, bias, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
    } else {
      if ((rc = pack_conv_weight(wp, dt, weight, Cout, Cin, k, Cout_r, Cin_p, st))) break;
      if (bias) e = cudaMemcpyAsync(bp, bias, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    if (e != cudaSuccess) { set_error("cudaMemcpyAsync: %s", cudaGetErrorString(e)); rc = BVG_ECUDA; break; }
    ConvArgs ca;
    ca.in = xin; ca.w = wp; ca.bias = bp; ca.out = yout; ca.res = res ? resp : nullptr;
    ca.accum = accum ? accp : nullptr; ca.scale = scale;
    ca.in_dtype = dt; ca.w_dtype = dt; ca.out_dtype = out_bf16 ? BVG_BF16 : BVG_F32;
    ca.B = B; ca.T = T; ca.Cin_p = Cin_p; ca.Cout_n = Cout_n; ca.Cout_r = Cout_r; ca.out_ld = Cout_n;
    ca.k = kk; ca.dil = up > 0 ? 1 : dil;
    if (act_taps) {
      // conv + bias (+ residual), then Activation1d: one kernel in bf16 mode (result rounded to bf16), two kernels otherwise
      ca.out_dtype = dt;
      if (dt == BVG_BF16 && !(variant & 16) && conv_act_fused_supported(ca)) {
        rc = conv_act_fused_launch(ca, alp, bep, *act_taps, st, res ? accp : nullptr);   // accp: y = conv + bias + res
        if (rc) break;
        rc = btc_to_bct(dst, yout, BVG_BF16, B, Cout, Cout_p, Tout, st);
        if (!rc && res) rc = btc_to_bct(dst_y, accp, BVG_F32, B, Cout, Cout_p, Tout, st);
        break;
      }
      // intermediate conv result: operand dtype without residual, the fp32 residual stream with it
      ca.out = accp;
      const int mid_dt = res ? BVG_F32 : dt;
      ca.out_dtype = mid_dt;
      if (dt == BVG_BF16 && conv_umma_supported(ca)) rc = conv_umma_launch(ca, variant, st);
      else rc = conv_simt_launch(ca, st);
      if (rc) break;
      rc = act1d_cl_launch(yout, accp, alp, bep, *act_taps, B, T, Cout_n, mid_dt, dt, dt == BVG_BF16, st);
      if (rc) break;
      rc = btc_to_bct(dst, yout, dt, B, Cout, Cout_p, Tout, st);
      if (!rc && res) rc = btc_to_bct(dst_y, accp, BVG_F32, B, Cout, Cout_p, Tout, st);
      break;
    }
    if (dt == BVG_BF16 && conv_umma_supported(ca)) rc = conv_umma_launch(ca, variant, st);
    else rc = conv_simt_launch(ca, st);
    if (rc) break;
    // [B, T, u*Cout_p] is [B, u*T, Cout_p]
    rc = btc_to_bct(dst, yout, out_bf16 ? BVG_BF16 : BVG_F32, B, Cout, Cout_p, Tout, st);
  } while (0);
  cudaFreeAsync(blk, st);
  return rc;
}

int bvg_conv1d_fwd(float* dst, const float* src, const float* weight, const float* bias, int B, int Cin, int Cout,
                   int64_t T, int k, int dilation, int mode, bvg_stream_t stream) {
  const int m = mode & 0xff;
  if (m != BVG_MODE_FP32 && m != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_conv1d_fwd: unknown mode %d", mode);
  return dense_layer(dst, src, weight, bias, B, Cin, Cout, T, k, dilation, 0, mode, (cudaStream_t)stream);
}

int bvg_conv1d_res_fwd(float* dst, const float* src, const float* weight, const float* bias, const float* res,
                       const float* accum, float scale, int out_bf16, int B, int Cin, int Cout, int64_t T, int k,
                       int dilation, int mode, bvg_stream_t stream) {
  const int m = mode & 0xff;
  if (m != BVG_MODE_FP32 && m != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_conv1d_res_fwd: unknown mode %d", mode);
  if (out_bf16 && m != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_conv1d_res_fwd: bf16 rounding of the result needs BVG_MODE_BF16");
  return dense_layer(dst, src, weight, bias, B, Cin, Cout, T, k, dilation, 0, mode, (cudaStream_t)stream, res, accum,
                     scale, out_bf16);
}

int bvg_conv1d_act_fwd(float* dst, const float* src, const float* weight, const float* bias, const float* alpha_log,
                       const float* beta_log, const float* up_taps, const float* down_taps, int B, int Cin, int Cout,
                       int64_t T, int k, int dilation, int mode, bvg_stream_t stream) {
  const int m = mode & 0xff;
  if (m != BVG_MODE_FP32 && m != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_conv1d_act_fwd: unknown mode %d", mode);
  if (!alpha_log || !beta_log || !up_taps || !down_taps) BVG_FAIL(BVG_EINVAL, "bvg_conv1d_act_fwd: null activation parameter");
  Taps taps;
  host_taps(&taps, up_taps, down_taps);
  return dense_layer(dst, src, weight, bias, B, Cin, Cout, T, k, dilation, 0, mode, (cudaStream_t)stream, nullptr, nullptr,
                     1.f, 0, alpha_log, beta_log, &taps);
}

int bvg_conv1d_res_act_fwd(float* dst_act, float* dst_y, const float* src, const float* weight, const float* bias,
                           const float* res, const float* alpha_log, const float* beta_log, const float* up_taps,
                           const float* down_taps, int B, int Cin, int Cout, int64_t T, int k, int dilation, int mode,
                           bvg_stream_t stream) {
  const int m = mode & 0xff;
  if (m != BVG_MODE_FP32 && m != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_conv1d_res_act_fwd: unknown mode %d", mode);
  if (!dst_y || !res || !alpha_log || !beta_log || !up_taps || !down_taps)
    BVG_FAIL(BVG_EINVAL, "bvg_conv1d_res_act_fwd: null pointer");
  Taps taps;
  host_taps(&taps, up_taps, down_taps);
  return dense_layer(dst_act, src, weight, bias, B, Cin, Cout, T, k, dilation, 0, mode, (cudaStream_t)stream, res, nullptr,
                     1.f, 0, alpha_log, beta_log, &taps, dst_y);
}

int bvg_convtr1d_fwd(float* dst, const float* src, const float* weight, const float* bias, int B, int Cin, int Cout,
                     int64_t T, int k, int stride, int mode, bvg_stream_t stream) {
  const int m = mode & 0xff;
  if (m != BVG_MODE_FP32 && m != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_convtr1d_fwd: unknown mode %d", mode);
  if (stride <= 0) BVG_FAIL(BVG_EINVAL, "bvg_convtr1d_fwd: bad stride");
  return dense_layer(dst, src, weight, bias, B, Cin, Cout, T, k, 1, stride, mode, (cudaStream_t)stream);
}

// One AMPBlock1 unit on the reference layout (tests, single-unit callers): transposes + packing around amp_unit.cu
// (BVG_MODE_BF16, one kernel) or around the layer-by-layer composition act -> conv -> act -> conv + residual.
int bvg_amp_unit_fwd(float* dst, const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* alpha1_log, const float* beta1_log, const float* alpha2_log, const float* beta2_log,
                     const float* up_taps, const float* down_taps, const float* accum, float scale, int out_bf16, int B,
                     int C, int64_t T, int k, int dilation, int mode, int flags, bvg_stream_t stream) {
  if (mode != BVG_MODE_FP32 && mode != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_amp_unit_fwd: unknown mode %d", mode);
  if (B < 0 || C <= 0 || T < 0 || k <= 0 || k % 2 != 1 || dilation < 1) BVG_FAIL(BVG_EINVAL, "bvg_amp_unit_fwd: bad dimension");
  if (out_bf16 && mode != BVG_MODE_BF16) BVG_FAIL(BVG_EINVAL, "bvg_amp_unit_fwd: bf16 rounding of the result needs BVG_MODE_BF16");
  if (B == 0 || T == 0) return BVG_OK;
  if (!dst || !x || !w1 || !w2 || !alpha1_log || !beta1_log || !alpha2_log || !beta2_log || !up_taps || !down_taps)
    BVG_FAIL(BVG_EINVAL, "bvg_amp_unit_fwd: null pointer");
  int rc = ensure_device_ok();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int dt = mode == BVG_MODE_BF16 ? BVG_BF16 : BVG_F32;
  const size_t es = dtype_size(dt);
  const int Cp = pad_channels(C, 16), Cr = round_up(Cp, 128);
  Taps taps;
  host_taps(&taps, up_taps, down_taps);
  const size_t a = 1024;
  auto up_a = [&](size_t v) { return (v + a - 1) / a * a; };
  const size_t b_f = up_a((size_t)B * T * Cp * 4), b_o = up_a((size_t)B * T * Cp * es), b_w = up_a((size_t)k * Cr * Cp * es),
               b_v = up_a((size_t)Cr * 4);
  unsigned char* blk = nullptr;
  BVG_CUDA(cudaMallocAsync((void**)&blk, 3 * b_f + 3 * b_o + 2 * b_w + 6 * b_v, st));
  unsigned char* q = blk;
  float* xin = (float*)q; q += b_f;
  float* yout = (float*)q; q += b_f;
  float* acc = (float*)q; q += b_f;
  void* t1 = q; q += b_o;
  void* t2 = q; q += b_o;
  void* t3 = q; q += b_o;
  void* wp1 = q; q += b_w;
  void* wp2 = q; q += b_w;
  float* vec = (float*)q;   // bias1, bias2, alpha1, beta1, alpha2, beta2 (zero padded)
  float* pv[6];
  for (int i = 0; i < 6; ++i) pv[i] = (float*)((unsigned char*)vec + i * b_v);
  const float* srcv[6] = {b1, b2, alpha1_log, beta1_log, alpha2_log, beta2_log};
  do {
    cudaError_t e = cudaMemsetAsync(vec, 0, 6 * b_v, st);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i)
      if (srcv[i]) e = cudaMemcpyAsync(pv[i], srcv[i], C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error("bvg_amp_unit_fwd: %s", cudaGetErrorString(e)); rc = BVG_ECUDA; break; }
    if ((rc = bct_to_btc(xin, BVG_F32, x, B, C, Cp, T, st))) break;
    if (accum && (rc = bct_to_btc(acc, BVG_F32, accum, B, C, Cp, T, st))) break;
    if ((rc = pack_conv_weight(wp1, dt, w1, C, C, k, Cr, Cp, st))) break;
    if ((rc = pack_conv_weight(wp2, dt, w2, C, C, k, Cr, Cp, st))) break;
    AmpUnitArgs ua;
    ua.x = xin; ua.out = yout; ua.accum = accum ? acc : nullptr; ua.scale = scale; ua.out_bf16 = out_bf16;
    ua.w1 = wp1; ua.w2 = wp2; ua.bias1 = pv[0]; ua.bias2 = pv[1];
    ua.al1 = pv[2]; ua.be1 = pv[3]; ua.al2 = pv[4]; ua.be2 = pv[5];
    ua.taps1 = taps; ua.taps2 = taps;
    ua.B = B; ua.T = T; ua.C = C; ua.Cp = Cp; ua.ld = Cp; ua.k = k; ua.dil = dilation;
    const bool fused = dt == BVG_BF16 && !(flags & 1) && amp_unit_supported(ua);
    if ((flags & 2) && !fused) { set_error("bvg_amp_unit_fwd: the one-kernel form does not take this unit"); rc = BVG_EINVAL; break; }
    if (fused) {
      if ((rc = amp_unit_launch(ua, st))) break;
    } else {
      // a1 -> c1 -> a2 -> c2 + residual, layer by layer
      const bool fast = dt == BVG_BF16;
      if ((rc = act1d_cl_launch(t1, xin, pv[2], pv[3], taps, B, T, Cp, BVG_F32, dt, fast, st))) break;
      ConvArgs ca;
      ca.in = t1; ca.w = wp1; ca.bias = pv[0]; ca.out = t2; ca.res = nullptr; ca.accum = nullptr; ca.scale = 1.f;
      ca.in_dtype = dt; ca.w_dtype = dt; ca.out_dtype = dt;
      ca.B = B; ca.T = T; ca.Cin_p = Cp; ca.Cout_n = Cp; ca.Cout_r = Cr; ca.out_ld = Cp; ca.k = k; ca.dil = dilation;
      rc = (dt == BVG_BF16 && conv_umma_supported(ca)) ? conv_umma_launch(ca, 0, st) : conv_simt_launch(ca, st);
      if (rc) break;
      if ((rc = act1d_cl_launch(t3, t2, pv[4], pv[5], taps, B, T, Cp, dt, dt, fast, st))) break;
      ca.in = t3; ca.w = wp2; ca.bias = pv[1]; ca.out = yout; ca.res = xin; ca.accum = accum ? acc : nullptr; ca.scale = scale;
      ca.out_dtype = out_bf16 ? BVG_BF16 : BVG_F32; ca.dil = 1;
      if (ca.accum && !ca.res) { rc = BVG_EINVAL; break; }
      rc = (dt == BVG_BF16 && conv_umma_supported(ca)) ? conv_umma_launch(ca, 0, st) : conv_simt_launch(ca, st);
      if (rc) break;
    }
    rc = btc_to_bct(dst, yout, out_bf16 ? BVG_BF16 : BVG_F32, B, C, Cp, T, st);
  } while (0);
  cudaFreeAsync(blk, st);
  return rc;
}

int bvg_create(const bvg_config* cfg, bvg_vocoder** out) { return vocoder_create(cfg, out); }
int bvg_set_tensor(bvg_vocoder* v, const char* name, const float* data, int64_t numel, int is_device) {
  return vocoder_set_tensor(v, name, data, numel, is_device);
}
int bvg_finalize(bvg_vocoder* v) { return vocoder_finalize(v); }
int64_t bvg_workspace_bytes(const bvg_vocoder* v, int B, int T0) { return vocoder_workspace_bytes(v, B, T0); }
int bvg_vocoder_fwd(bvg_vocoder* v, const float* mel, float* wav, int B, int T0, bvg_stream_t stream) {
  return vocoder_forward(v, mel, nullptr, wav, 0, B, T0, (cudaStream_t)stream);
}
int bvg_vocoder_fwd_cond(bvg_vocoder* v, const float* latent, const float* spk_emb, float* wav, int B, int T0,
                         bvg_stream_t stream) {
  if (!spk_emb) BVG_FAIL(BVG_EINVAL, "bvg_vocoder_fwd_cond: null speaker embedding");
  return vocoder_forward(v, latent, spk_emb, wav, 0, B, T0, (cudaStream_t)stream);
}
int bvg_vocoder_fwd_host(bvg_vocoder* v, const float* mel_host, void* wav_host, int wav_dtype, int B, int T0,
                         bvg_stream_t stream) {
  return vocoder_forward_host(v, mel_host, wav_host, wav_dtype, B, T0, (cudaStream_t)stream);
}

}  // extern "C"
