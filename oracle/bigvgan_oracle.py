"""ORACLE — test infrastructure, not product code.

CPU restatement (torch tensors, any float dtype incl. float64) of the BigVGAN v2
generator path of caishiqing/voice-tts.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module;
the product package never does.

Pinned against the reference: the reference has no tests or golden vectors of
its own, so parity is pinned on outputs of the reference
itself, imported in the build container by `oracle/make_golden.py` and committed
under `tests/golden/` (activation, AMP block, tiny generator, full-size
generator slices, filter taps).  `tests/test_oracle.py` checks this file against
every one of them.

What each function follows (paths relative to the reference root):
  kaiser_taps            indextts/s2mel/modules/bigvgan/alias_free_activation/torch/filter.py:30-62
  upsample2x             .../alias_free_activation/torch/resample.py:29-38   (closed polyphase form)
  snake / snakebeta      indextts/s2mel/modules/bigvgan/activations.py:46-59,107-120
  downsample2x           .../alias_free_activation/torch/filter.py:94-101, resample.py:55-58
  activation1d           .../alias_free_activation/torch/act.py:25-30
  conv1d / conv_transpose1d   torch.nn.Conv1d / ConvTranspose1d as built at bigvgan.py:59-66,76-83,285-287,306-312,348-350
  amp_block1             indextts/s2mel/modules/bigvgan/bigvgan.py:132-141
  generator_forward      indextts/s2mel/modules/bigvgan/bigvgan.py:360-386
  generator_v1_forward   indextts/BigVGAN/models.py:212-250 (the IndexTTS-v1 speaker-conditioned generator, downstream of
                         its ECAPA speaker encoder: the embedding is an input)
  fold_weight_norm       torch.nn.utils.weight_norm semantics used at bigvgan.py:388-400

The FIR stages are written in their closed polyphase form with explicit index
clamps instead of pad/conv_transpose/slice, so this file is an independent
statement of the arithmetic; `activation1d_staged` is the literal
pad -> transposed-conv -> slice -> act -> pad -> strided-conv sequence and the
tests require both to agree.
"""
import math

import torch
import torch.nn.functional as F

NO_DIV_BY_ZERO = 1e-9


# --------------------------------------------------------------------------
# filter design
# --------------------------------------------------------------------------
def kaiser_taps(cutoff=0.25, half_width=0.3, n=12, dtype=torch.float32):
    half = n // 2
    att = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    if att > 50.0:
        beta = 0.1102 * (att - 8.7)
    elif att >= 21.0:
        beta = 0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0)
    else:
        beta = 0.0
    w = torch.kaiser_window(n, beta=beta, periodic=False, dtype=dtype)
    t = (torch.arange(-half, half, dtype=dtype) + 0.5) if n % 2 == 0 else (torch.arange(n, dtype=dtype) - half)
    taps = 2 * cutoff * w * torch.sinc(2 * cutoff * t)
    return taps / taps.sum()


# --------------------------------------------------------------------------
# anti-aliased activation, closed form
# --------------------------------------------------------------------------
def _gather_clamped(x, idx):
    """x[..., clamp(idx, 0, T-1)] for an index tensor idx of any shape."""
    T = x.shape[-1]
    return x[..., idx.clamp(0, T - 1)]


def upsample2x(x, taps):
    """u[2t]   = 2*sum_q taps[2q+1]*x[clamp(t+2-q)]
       u[2t+1] = 2*sum_q taps[2q]  *x[clamp(t+3-q)],  q = 0..5, clamp to [0,T-1]."""
    T = x.shape[-1]
    t = torch.arange(T)
    even = torch.zeros_like(x)
    odd = torch.zeros_like(x)
    for q in range(6):
        even = even + taps[2 * q + 1] * _gather_clamped(x, t + 2 - q)
        odd = odd + taps[2 * q] * _gather_clamped(x, t + 3 - q)
    u = torch.stack((2 * even, 2 * odd), dim=-1)
    return u.reshape(*x.shape[:-1], 2 * T)


def snakebeta(u, alpha, beta, logscale=True):
    """u + 1/(b+1e-9) * sin(a*u)^2 with per-channel a,b (exp() of the stored
    parameters when logscale).  u is [B,C,T]."""
    a = alpha.reshape(1, -1, 1).to(u.dtype)
    b = beta.reshape(1, -1, 1).to(u.dtype)
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    return u + (1.0 / (b + NO_DIV_BY_ZERO)) * torch.sin(u * a) ** 2


def downsample2x(v, taps):
    """y[t] = sum_k taps[k]*v[clamp(2t+k-5, 0, 2T-1)], k = 0..11."""
    T = v.shape[-1] // 2
    t = torch.arange(T)
    y = torch.zeros(*v.shape[:-1], T, dtype=v.dtype)
    for k in range(12):
        y = y + taps[k] * _gather_clamped(v, 2 * t + k - 5)
    return y


def activation1d(x, alpha, beta, up_taps, down_taps, logscale=True):
    up_taps = up_taps.reshape(-1).to(x.dtype)
    down_taps = down_taps.reshape(-1).to(x.dtype)
    return downsample2x(snakebeta(upsample2x(x, up_taps), alpha, beta, logscale), down_taps)


def activation1d_staged(x, alpha, beta, up_taps, down_taps, logscale=True):
    """Same operator as the literal sequence of padded depthwise convolutions."""
    C = x.shape[1]
    fu = up_taps.reshape(1, 1, 12).to(x.dtype).expand(C, -1, -1)
    fd = down_taps.reshape(1, 1, 12).to(x.dtype).expand(C, -1, -1)
    u = 2 * F.conv_transpose1d(F.pad(x, (5, 5), mode="replicate"), fu, stride=2, groups=C)[..., 15:-15]
    v = snakebeta(u, alpha, beta, logscale)
    return F.conv1d(F.pad(v, (5, 6), mode="replicate"), fd, stride=2, groups=C)


# --------------------------------------------------------------------------
# dense layers
# --------------------------------------------------------------------------
def conv1d(x, w, b, dilation=1):
    k = w.shape[-1]
    return F.conv1d(x, w.to(x.dtype), None if b is None else b.to(x.dtype),
                    padding=(k - 1) * dilation // 2, dilation=dilation)


def conv_transpose1d(x, w, b, stride):
    k = w.shape[-1]
    return F.conv_transpose1d(x, w.to(x.dtype), None if b is None else b.to(x.dtype),
                              stride=stride, padding=(k - stride) // 2)


def conv1d_indexed(x, w, b, dilation=1):
    """out[co,t] = b[co] + sum_j W[co,:,j] . x[:, t+(j-(k-1)/2)*d], zero outside
    (SURVEY.md 8(a)); slow, used by tests to pin `conv1d`."""
    B, C, T = x.shape
    k = w.shape[-1]
    out = torch.zeros(B, w.shape[0], T, dtype=x.dtype)
    for j in range(k):
        s = (j - (k - 1) // 2) * dilation
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            out[:, :, lo:hi] += torch.einsum("oc,bct->bot", w[:, :, j].to(x.dtype), x[:, :, lo + s:hi + s])
    if b is not None:
        out += b.to(x.dtype).view(1, -1, 1)
    return out


def conv_transpose1d_polyphase(x, w, b, stride):
    """Polyphase statement of ConvTranspose1d(k=2u, stride=u, pad=u/2):
    for t=u*m+r, s=r+u/2:  s<u : W[:,:,s]^T x[m] + W[:,:,s+u]^T x[m-1]
                           else: W[:,:,s-u]^T x[m+1] + W[:,:,s]^T x[m]."""
    B, Cin, T = x.shape
    u = stride
    Cout = w.shape[1]
    assert w.shape[-1] == 2 * u
    wd = w.to(x.dtype)
    xp = F.pad(x, (1, 1))
    out = torch.zeros(B, Cout, T, u, dtype=x.dtype)
    for r in range(u):
        s = r + u // 2
        if s < u:
            out[..., r] = torch.einsum("co,bct->bot", wd[:, :, s], xp[:, :, 1:T + 1]) + \
                          torch.einsum("co,bct->bot", wd[:, :, s + u], xp[:, :, 0:T])
        else:
            out[..., r] = torch.einsum("co,bct->bot", wd[:, :, s - u], xp[:, :, 2:T + 2]) + \
                          torch.einsum("co,bct->bot", wd[:, :, s], xp[:, :, 1:T + 1])
    out = out.reshape(B, Cout, T * u)
    if b is not None:
        out = out + b.to(x.dtype).view(1, -1, 1)
    return out


def fold_weight_norm(sd):
    """weight = g * v / ||v||, norm over all dims but 0 (legacy
    torch.nn.utils.weight_norm, dim=0).  Returns a new dict without *_g/*_v."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            g = sd[k[:-2] + "_g"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            out[k[:-9] + ".weight"] = v * (g / norm)
        elif k.endswith(".weight_g"):
            continue
        else:
            out[k] = v
    return out


# --------------------------------------------------------------------------
# blocks and generator
# --------------------------------------------------------------------------
_STAGED = False   # see `staged_ops`


class staged_ops:
    """Context manager: inside it every Activation1d of the generator runs as `activation1d_staged`, i.e. the reference's
    own operator sequence (F.pad replicate -> conv_transpose1d -> slice -> snake -> F.pad -> strided conv1d, resample.py:29-38,
    filter.py:94-101) instead of the closed polyphase form.  The two agree to rounding (tests/test_oracle.py); the staged form
    is what `bench.py` times as the CPU arm, because it costs what the reference costs on the host (measured in the build
    container against the imported reference)."""

    def __enter__(self):
        global _STAGED
        self.prev, _STAGED = _STAGED, True

    def __exit__(self, *exc):
        global _STAGED
        _STAGED = self.prev


def _act(sd, prefix, x, h):
    logscale = h.get("snake_logscale", True)
    alpha = sd[prefix + ".act.alpha"]
    beta = sd[prefix + ".act.beta"] if h["activation"] == "snakebeta" else alpha
    fn = activation1d_staged if _STAGED else activation1d
    return fn(x, alpha, beta, sd[prefix + ".upsample.filter"], sd[prefix + ".downsample.lowpass.filter"], logscale)


def amp_block1(sd, prefix, x, h, dilations):
    for l, d in enumerate(dilations):
        xt = _act(sd, "%s.activations.%d" % (prefix, 2 * l), x, h)
        xt = conv1d(xt, sd["%s.convs1.%d.weight" % (prefix, l)], sd["%s.convs1.%d.bias" % (prefix, l)], d)
        xt = _act(sd, "%s.activations.%d" % (prefix, 2 * l + 1), xt, h)
        xt = conv1d(xt, sd["%s.convs2.%d.weight" % (prefix, l)], sd["%s.convs2.%d.bias" % (prefix, l)], 1)
        x = xt + x
    return x


def generator_forward(sd, h, mel, dtype=None, taps=None):
    """mel [B,num_mels,T] -> wav [B,1,T*prod(upsample_rates)].  `sd` is a folded
    or weight-normed state dict in the reference's key names."""
    if any(k.endswith(".weight_v") for k in sd):
        sd = fold_weight_norm(sd)
    if dtype is not None:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        mel = mel.to(dtype)
    if h.get("resblock", "1") != "1":
        raise NotImplementedError("only AMPBlock1 is live in the reference (AMPBlock2.forward returns None)")
    nk = len(h["resblock_kernel_sizes"])
    x = conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"])
    for i, u in enumerate(h["upsample_rates"]):
        x = conv_transpose1d(x, sd["ups.%d.0.weight" % i], sd["ups.%d.0.bias" % i], u)
        xs = None
        for j in range(nk):
            y = amp_block1(sd, "resblocks.%d" % (i * nk + j), x, h, h["resblock_dilation_sizes"][j])
            xs = y if xs is None else xs + y
        x = xs / nk
    x = _act(sd, "activation_post", x, h)
    x = conv1d(x, sd["conv_post.weight"], sd.get("conv_post.bias"))
    if h.get("use_tanh_at_final", True):
        return torch.tanh(x)
    return x.clamp(-1.0, 1.0)


def conv_transpose1d_taps(x, w, b, stride):
    """General 3-tap polyphase statement of ConvTranspose1d(k, stride=u, pad=(k-u)/2), k-u even, pad <= u:
    out[u*m + r] = sum_{tap in 0..2} W[:, :, r + pad + u*(1 - tap)]^T x[m + tap - 1]  (index in [0, k), x zero outside).
    Used by the tests to pin the packing of the v1 plan (k = u and k = 2u layers)."""
    B, Cin, T = x.shape
    u = stride
    k = w.shape[-1]
    pad = (k - u) // 2
    assert (k - u) % 2 == 0 and 0 <= pad <= u
    wd = w.to(x.dtype)
    xp = F.pad(x, (1, 1))
    out = torch.zeros(B, w.shape[1], T, u, dtype=x.dtype)
    for r in range(u):
        for tap in range(3):
            j = r + pad + u * (1 - tap)
            if 0 <= j < k:
                out[..., r] += torch.einsum("co,bct->bot", wd[:, :, j], xp[:, :, tap:tap + T])
    out = out.reshape(B, w.shape[1], T * u)
    if b is not None:
        out = out + b.to(x.dtype).view(1, -1, 1)
    return out


def generator_v1_forward(sd, h, latent, spk_emb, dtype=None):
    """IndexTTS-v1 generator (indextts/BigVGAN/models.py:212-250) after its speaker encoder:
    latent [B, T, gpt_dim] (feat_upsample = False: `x.transpose(1, 2)`, :220), spk_emb [B, E] (the `[B, 1, E]` output of
    `self.speaker_encoder`, transposed to [B, E, 1] at :214) -> wav [B, 1, T*prod(upsample_rates)].
      x = conv_pre(x) + cond_layer(e)                      :223-224
      per stage: x = ups[i](x) (+ conds[i](e))             :228-234, then the mean of the AMP blocks :236-243
      x = tanh(conv_post(activation_post(x)))              :246-248"""
    if any(k.endswith(".weight_v") for k in sd):
        sd = fold_weight_norm(sd)
    if dtype is not None:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        latent = latent.to(dtype)
        spk_emb = spk_emb.to(dtype)
    if h.get("feat_upsample", False):
        raise NotImplementedError("feat_upsample=True (4x linear interpolation of the latent) is not on the shipped v1 path")
    nk = len(h["resblock_kernel_sizes"])
    e = spk_emb.unsqueeze(-1)                                         # [B, E, 1]
    x = conv1d(latent.transpose(1, 2), sd["conv_pre.weight"], sd["conv_pre.bias"])
    x = x + conv1d(e, sd["cond_layer.weight"], sd["cond_layer.bias"])
    for i, u in enumerate(h["upsample_rates"]):
        x = conv_transpose1d(x, sd["ups.%d.0.weight" % i], sd["ups.%d.0.bias" % i], u)
        if h.get("cond_d_vector_in_each_upsampling_layer", False):
            x = x + conv1d(e, sd["conds.%d.weight" % i], sd["conds.%d.bias" % i])
        xs = None
        for j in range(nk):
            y = amp_block1(sd, "resblocks.%d" % (i * nk + j), x, h, h["resblock_dilation_sizes"][j])
            xs = y if xs is None else xs + y
        x = xs / nk
    x = _act(sd, "activation_post", x, h)
    x = conv1d(x, sd["conv_post.weight"], sd.get("conv_post.bias"))
    return torch.tanh(x)


def snr_db(ref, test):
    ref = ref.double().flatten()
    err = test.double().flatten() - ref
    return float(10 * torch.log10(ref.pow(2).sum() / err.pow(2).sum().clamp_min(1e-300)))
