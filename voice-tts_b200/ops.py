"""torch custom ops over the C ABI (plumbing only: device memory, streams).

  bvg_b200::act1d        fused anti-aliased Snake/SnakeBeta on [B,C,T]   (reference:
                         anti_alias_activation_cuda.forward, cuda/activation1d.py:21-27)
  bvg_b200::act1d_cl     same on channels-last [B,T,C]
  bvg_b200::conv1d / conv_transpose1d   single dense layers (tests, single-layer callers)
  bvg_b200::vocoder      whole generator through a native handle
"""
from typing import List

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
_DT_ACT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}   # bvg_act1d_fwd also takes half


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


@torch.library.custom_op("bvg_b200::act1d", mutates_args=())
def act1d(x: torch.Tensor, alpha_log: torch.Tensor, beta_log: torch.Tensor, up_taps: List[float],
          down_taps: List[float], fast: bool) -> torch.Tensor:
    if x.dim() != 3:
        raise RuntimeError("act1d expects [B, C, T], got %s" % (tuple(x.shape),))
    _require_cuda(x, "x")
    if x.dtype not in _DT_ACT:
        raise RuntimeError("act1d supports float32, bfloat16 and float16, got %s" % x.dtype)
    B, C, T = x.shape
    a = alpha_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b = beta_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if a.numel() != C or b.numel() != C:
        raise RuntimeError("alpha/beta must have C=%d elements" % C)
    y = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.bvg_act1d_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), _lib.taps_array(up_taps),
                               _lib.taps_array(down_taps), B, C, T, _DT_ACT[x.dtype],
                               _lib.ACT_FAST_SIN if fast else 0, _stream(x))
    _lib.check(rc, "bvg_act1d_fwd")
    return y


@act1d.register_fake
def _(x, alpha_log, beta_log, up_taps, down_taps, fast):
    return torch.empty_like(x)


@torch.library.custom_op("bvg_b200::act1d_cl", mutates_args=())
def act1d_cl(x: torch.Tensor, alpha_log: torch.Tensor, beta_log: torch.Tensor, up_taps: List[float],
             down_taps: List[float], out_bf16: bool, fast: bool) -> torch.Tensor:
    if x.dim() != 3:
        raise RuntimeError("act1d_cl expects [B, T, C]")
    _require_cuda(x, "x")
    if x.dtype not in _DT:
        raise RuntimeError("act1d_cl supports float32 and bfloat16, got %s" % x.dtype)
    B, T, C = x.shape
    a = alpha_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b = beta_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.bvg_act1d_cl_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), _lib.taps_array(up_taps),
                                  _lib.taps_array(down_taps), B, T, C, _DT[x.dtype], _DT[y.dtype],
                                  _lib.ACT_FAST_SIN if fast else 0, _stream(x))
    _lib.check(rc, "bvg_act1d_cl_fwd")
    return y


@act1d_cl.register_fake
def _(x, alpha_log, beta_log, up_taps, down_taps, out_bf16, fast):
    return torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)


def _mode(precision, variant=0):
    m = {"fp32": _lib.MODE_FP32, "bf16": _lib.MODE_BF16}.get(precision)
    if m is None:
        raise RuntimeError("precision must be 'fp32' or 'bf16', got %r" % (precision,))
    return m | (int(variant) << 8)


@torch.library.custom_op("bvg_b200::conv1d", mutates_args=())
def conv1d(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, dilation: int, precision: str,
           variant: int) -> torch.Tensor:
    _require_cuda(x, "x")
    B, Cin, T = x.shape
    Cout, Cin2, k = weight.shape
    if Cin2 != Cin or x.dtype != torch.float32:
        raise RuntimeError("conv1d: expects fp32 [B,Cin,T] and weight [Cout,Cin,k]")
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    bptr = 0
    if bias.numel():
        bb = bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
        bptr = bb.data_ptr()
    y = torch.empty(B, Cout, T, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_conv1d_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), bptr, B, Cin, Cout, T, k, dilation,
                                        _mode(precision, variant), _stream(x))
    _lib.check(rc, "bvg_conv1d_fwd")
    return y


@conv1d.register_fake
def _(x, weight, bias, dilation, precision, variant):
    return x.new_empty(x.shape[0], weight.shape[0], x.shape[2])


@torch.library.custom_op("bvg_b200::conv1d_res", mutates_args=())
def conv1d_res(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, res: torch.Tensor, accum: torch.Tensor,
               scale: float, out_bf16: bool, dilation: int, precision: str, variant: int) -> torch.Tensor:
    """(conv1d(x) + bias + res) * scale + accum; empty tensors stand for absent operands
    (AMPBlock1 residual bigvgan.py:132-141 and resblock mean :369-375 folded into the conv epilogue)."""
    _require_cuda(x, "x")
    B, Cin, T = x.shape
    Cout, Cin2, k = weight.shape
    if Cin2 != Cin or x.dtype != torch.float32:
        raise RuntimeError("conv1d_res: expects fp32 [B,Cin,T] and weight [Cout,Cin,k]")
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    keep = []

    def ptr(t, shape):
        if not t.numel():
            return 0
        if tuple(t.shape) != shape:
            raise RuntimeError("conv1d_res: operand shape %s, expected %s" % (tuple(t.shape), shape))
        tt = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        keep.append(tt)
        return tt.data_ptr()

    bptr, rptr, aptr = ptr(bias, (Cout,)), ptr(res, (B, Cout, T)), ptr(accum, (B, Cout, T))
    y = torch.empty(B, Cout, T, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_conv1d_res_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), bptr, rptr, aptr, float(scale),
                                            int(out_bf16), B, Cin, Cout, T, k, dilation, _mode(precision, variant),
                                            _stream(x))
    _lib.check(rc, "bvg_conv1d_res_fwd")
    return y


@conv1d_res.register_fake
def _(x, weight, bias, res, accum, scale, out_bf16, dilation, precision, variant):
    return x.new_empty(x.shape[0], weight.shape[0], x.shape[2])


@torch.library.custom_op("bvg_b200::conv1d_act", mutates_args=())
def conv1d_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, alpha_log: torch.Tensor,
               beta_log: torch.Tensor, up_taps: List[float], down_taps: List[float], dilation: int, precision: str,
               variant: int) -> torch.Tensor:
    """Activation1d(conv1d(x) + bias): `xt = c1(xt); xt = a2(xt)` of AMPBlock1.forward (bigvgan.py:136-138) as one
    tcgen05 kernel in bf16 mode (variant 16 forces the two-kernel composition)."""
    _require_cuda(x, "x")
    B, Cin, T = x.shape
    Cout, Cin2, k = weight.shape
    if Cin2 != Cin or x.dtype != torch.float32:
        raise RuntimeError("conv1d_act: expects fp32 [B,Cin,T] and weight [Cout,Cin,k]")
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    a = alpha_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b = beta_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if a.numel() != Cout or b.numel() != Cout:
        raise RuntimeError("conv1d_act: alpha/beta must have Cout=%d elements" % Cout)
    bptr = 0
    if bias.numel():
        bb = bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
        bptr = bb.data_ptr()
    y = torch.empty(B, Cout, T, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_conv1d_act_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), bptr, a.data_ptr(), b.data_ptr(),
                                            _lib.taps_array(up_taps), _lib.taps_array(down_taps), B, Cin, Cout, T, k,
                                            dilation, _mode(precision, variant), _stream(x))
    _lib.check(rc, "bvg_conv1d_act_fwd")
    return y


@conv1d_act.register_fake
def _(x, weight, bias, alpha_log, beta_log, up_taps, down_taps, dilation, precision, variant):
    return x.new_empty(x.shape[0], weight.shape[0], x.shape[2])


@torch.library.custom_op("bvg_b200::conv1d_res_act", mutates_args=())
def conv1d_res_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, res: torch.Tensor, alpha_log: torch.Tensor,
                   beta_log: torch.Tensor, up_taps: List[float], down_taps: List[float], dilation: int, precision: str,
                   variant: int) -> List[torch.Tensor]:
    """[Activation1d(y), y] with y = conv1d(x) + bias + res: `xt = c2(xt); x = xt + x` and the next unit's `a1(x)`
    (bigvgan.py:134-139) as one tcgen05 kernel in bf16 mode (variant 16 forces the two-kernel composition)."""
    _require_cuda(x, "x")
    B, Cin, T = x.shape
    Cout, Cin2, k = weight.shape
    if Cin2 != Cin or x.dtype != torch.float32 or tuple(res.shape) != (B, Cout, T):
        raise RuntimeError("conv1d_res_act: expects fp32 x [B,Cin,T], weight [Cout,Cin,k], res [B,Cout,T]")
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    a = alpha_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b = beta_log.detach().to(device=x.device, dtype=torch.float32).contiguous()
    r = res.detach().to(device=x.device, dtype=torch.float32).contiguous()
    bptr = 0
    if bias.numel():
        bb = bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
        bptr = bb.data_ptr()
    ya = torch.empty(B, Cout, T, device=x.device, dtype=torch.float32)
    y = torch.empty(B, Cout, T, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_conv1d_res_act_fwd(ya.data_ptr(), y.data_ptr(), x.data_ptr(), w.data_ptr(), bptr, r.data_ptr(),
                                                a.data_ptr(), b.data_ptr(), _lib.taps_array(up_taps),
                                                _lib.taps_array(down_taps), B, Cin, Cout, T, k, dilation,
                                                _mode(precision, variant), _stream(x))
    _lib.check(rc, "bvg_conv1d_res_act_fwd")
    return [ya, y]


@conv1d_res_act.register_fake
def _(x, weight, bias, res, alpha_log, beta_log, up_taps, down_taps, dilation, precision, variant):
    return [x.new_empty(x.shape[0], weight.shape[0], x.shape[2]), x.new_empty(x.shape[0], weight.shape[0], x.shape[2])]


@torch.library.custom_op("bvg_b200::amp_unit", mutates_args=())
def amp_unit(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
             alpha1_log: torch.Tensor, beta1_log: torch.Tensor, alpha2_log: torch.Tensor, beta2_log: torch.Tensor,
             up_taps: List[float], down_taps: List[float], accum: torch.Tensor, scale: float, out_bf16: bool,
             dilation: int, precision: str, form: int) -> torch.Tensor:
    """One iteration of AMPBlock1.forward (bigvgan.py:132-141): (x + c2(a2(c1(a1(x))))) * scale + accum on fp32 [B, C, T].
    form: 0 = one kernel where the unit qualifies, 1 = layer by layer, 2 = one kernel or fail."""
    _require_cuda(x, "x")
    B, C, T = x.shape
    k = w1.shape[-1]
    if x.dtype != torch.float32 or tuple(w1.shape) != (C, C, k) or tuple(w2.shape) != (C, C, k):
        raise RuntimeError("amp_unit: expects fp32 [B,C,T] and two weights [C,C,k]")
    keep = []

    def ptr(t, n):
        if not t.numel():
            return 0
        if t.numel() != n:
            raise RuntimeError("amp_unit: operand has %d elements, expected %d" % (t.numel(), n))
        tt = t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        keep.append(tt)
        return tt.data_ptr()

    args = [ptr(w1, C * C * k), ptr(b1, C), ptr(w2, C * C * k), ptr(b2, C), ptr(alpha1_log, C), ptr(beta1_log, C),
            ptr(alpha2_log, C), ptr(beta2_log, C)]
    aptr = ptr(accum, B * C * T)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_amp_unit_fwd(y.data_ptr(), x.data_ptr(), *args, _lib.taps_array(up_taps),
                                          _lib.taps_array(down_taps), aptr, float(scale), int(out_bf16), B, C, T, k,
                                          dilation, _mode(precision), int(form), _stream(x))
    _lib.check(rc, "bvg_amp_unit_fwd")
    return y


@amp_unit.register_fake
def _(x, w1, b1, w2, b2, alpha1_log, beta1_log, alpha2_log, beta2_log, up_taps, down_taps, accum, scale, out_bf16,
      dilation, precision, form):
    return torch.empty_like(x)


@torch.library.custom_op("bvg_b200::conv_transpose1d", mutates_args=())
def conv_transpose1d(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, stride: int, precision: str,
                     variant: int) -> torch.Tensor:
    _require_cuda(x, "x")
    B, Cin, T = x.shape
    Cin2, Cout, k = weight.shape
    if Cin2 != Cin or x.dtype != torch.float32:
        raise RuntimeError("conv_transpose1d: expects fp32 [B,Cin,T] and weight [Cin,Cout,k]")
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    bptr = 0
    if bias.numel():
        bb = bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
        bptr = bb.data_ptr()
    y = torch.empty(B, Cout, T * stride, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_convtr1d_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), bptr, B, Cin, Cout, T, k, stride,
                                          _mode(precision, variant), _stream(x))
    _lib.check(rc, "bvg_convtr1d_fwd")
    return y


@conv_transpose1d.register_fake
def _(x, weight, bias, stride, precision, variant):
    return x.new_empty(x.shape[0], weight.shape[1], x.shape[2] * stride)


# ---- whole generator -----------------------------------------------------------------------------
_HANDLES = {}  # id -> (ctypes handle, num_mels, total_upsample, device index)


def register_handle(handle, num_mels, total_up, device_index):
    hid = max(_HANDLES.keys(), default=0) + 1
    _HANDLES[hid] = (handle, num_mels, total_up, device_index)
    return hid


def release_handle(hid):
    ent = _HANDLES.pop(hid, None)
    if ent is not None:
        _lib.load().bvg_destroy(ent[0])


@torch.library.custom_op("bvg_b200::vocoder", mutates_args=())
def vocoder(mel: torch.Tensor, handle_id: int) -> torch.Tensor:
    handle, num_mels, total_up, dev = _HANDLES[handle_id]
    _require_cuda(mel, "mel")
    if mel.dim() != 3 or mel.shape[1] != num_mels:
        raise RuntimeError("vocoder expects mel [B, %d, T], got %s" % (num_mels, tuple(mel.shape)))
    if mel.dtype != torch.float32:
        raise RuntimeError("vocoder expects a float32 mel (as infer_v2.py:735 passes), got %s" % mel.dtype)
    if mel.device.index != dev:
        raise RuntimeError("mel is on cuda:%s but the vocoder handle lives on cuda:%d" % (mel.device.index, dev))
    B, _, T = mel.shape
    wav = torch.empty(B, 1, T * total_up, device=mel.device, dtype=torch.float32)
    with torch.cuda.device(mel.device):
        rc = _lib.load().bvg_vocoder_fwd(handle, mel.data_ptr(), wav.data_ptr(), B, T, _stream(mel))
    _lib.check(rc, "bvg_vocoder_fwd")
    return wav


@vocoder.register_fake
def _(mel, handle_id):
    _, _, total_up, _ = _HANDLES[handle_id]
    return mel.new_empty(mel.shape[0], 1, mel.shape[2] * total_up)


@torch.library.custom_op("bvg_b200::vocoder_cond", mutates_args=())
def vocoder_cond(latent: torch.Tensor, spk_emb: torch.Tensor, handle_id: int) -> torch.Tensor:
    """speaker-conditioned v1 generator (indextts/BigVGAN/models.py:212-250): latent [B, T, gpt_dim], spk_emb [B, E]."""
    handle, cin, total_up, dev = _HANDLES[handle_id]
    _require_cuda(latent, "latent")
    _require_cuda(spk_emb, "spk_emb")
    if latent.dim() != 3 or latent.shape[2] != cin:
        raise RuntimeError("vocoder_cond expects latent [B, T, %d], got %s" % (cin, tuple(latent.shape)))
    if latent.dtype != torch.float32 or spk_emb.dtype != torch.float32:
        raise RuntimeError("vocoder_cond expects float32 tensors")
    if spk_emb.dim() != 2 or spk_emb.shape[0] != latent.shape[0]:
        raise RuntimeError("vocoder_cond expects spk_emb [B, E], got %s" % (tuple(spk_emb.shape),))
    if latent.device.index != dev or spk_emb.device.index != dev:
        raise RuntimeError("inputs are not on the vocoder handle's device cuda:%d" % dev)
    B, T, _ = latent.shape
    wav = torch.empty(B, 1, T * total_up, device=latent.device, dtype=torch.float32)
    with torch.cuda.device(latent.device):
        rc = _lib.load().bvg_vocoder_fwd_cond(handle, latent.data_ptr(), spk_emb.data_ptr(), wav.data_ptr(), B, T,
                                              _stream(latent))
    _lib.check(rc, "bvg_vocoder_fwd_cond")
    return wav


@vocoder_cond.register_fake
def _(latent, spk_emb, handle_id):
    _, _, total_up, _ = _HANDLES[handle_id]
    return latent.new_empty(latent.shape[0], 1, latent.shape[1] * total_up)
