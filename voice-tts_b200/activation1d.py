"""Drop-in `Activation1d` / `Snake` / `SnakeBeta` with the reference's attribute
and state-dict layout (`.act.alpha`, `.act.beta`, `.upsample.filter`,
`.downsample.lowpass.filter`), executed by ONE fused sm_100a kernel.

Mirrors: alias_free_activation/cuda/activation1d.py:35-77 (constructor signature,
log-scale handling, Snake -> beta == alpha), torch/resample.py:10-58 and
torch/filter.py:65-101 (buffer-holding modules), activations.py:9-120.
Unlike the reference's CUDA kernel the result equals the torch operator at the
sequence edges too.  Forward only, CUDA only.
"""
import torch
import torch.nn as nn

from . import ops
from .synth import kaiser_sinc_filter1d


class Snake(nn.Module):
    """x + 1/a * sin^2(a x), per-channel a (exp(a) when alpha_logscale)."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(init * alpha, requires_grad=alpha_trainable)
        self.no_div_by_zero = 1e-9


class SnakeBeta(nn.Module):
    """x + 1/b * sin^2(a x), per-channel a, b (exp() when alpha_logscale)."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(init * alpha, requires_grad=alpha_trainable)
        self.beta = nn.Parameter(init.clone() * alpha, requires_grad=alpha_trainable)
        self.no_div_by_zero = 1e-9


class LowPassFilter1d(nn.Module):
    def __init__(self, cutoff=0.5, half_width=0.6, stride=1, padding=True, padding_mode="replicate", kernel_size=12):
        super().__init__()
        if cutoff < 0.0:
            raise ValueError("Minimum cutoff must be larger than zero.")
        if cutoff > 0.5:
            raise ValueError("A cutoff above 0.5 does not make sense.")
        self.kernel_size = kernel_size
        self.stride = stride
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff, half_width, kernel_size))


class UpSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.register_buffer("filter", kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, self.kernel_size))


class DownSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.lowpass = LowPassFilter1d(0.5 / ratio, 0.6 / ratio, stride=ratio, kernel_size=self.kernel_size)


class Activation1d(nn.Module):
    """up x2 -> Snake/SnakeBeta -> down x2 as one kernel; x: [B, C, T] -> [B, C, T]."""

    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2, up_kernel_size: int = 12,
                 down_kernel_size: int = 12, fused: bool = True, fast_sin=None):
        super().__init__()
        if up_ratio != 2 or down_ratio != 2 or up_kernel_size != 12 or down_kernel_size != 12:
            # the reference's fused kernel silently computes the wrong thing here
            # (cuda/activation1d.py:16-18); refuse instead
            raise NotImplementedError("the fused kernel implements ratio 2 / 12-tap filters only")
        self.up_ratio, self.down_ratio = up_ratio, down_ratio
        self.act = activation
        self.upsample = UpSample1d(up_ratio, up_kernel_size)
        self.downsample = DownSample1d(down_ratio, down_kernel_size)
        self.fused = fused          # kept for signature compatibility; there is only the fused path
        self.fast_sin = fast_sin    # None: fast for bf16 inputs, accurate for fp32
        self._taps_key = None
        self._taps = None

    def _host_taps(self):
        fu, fd = self.upsample.filter, self.downsample.lowpass.filter
        key = (fu._version, fd._version, fu.data_ptr(), fd.data_ptr())
        if key != self._taps_key:
            self._taps = (fu.detach().reshape(-1).float().cpu().tolist(), fd.detach().reshape(-1).float().cpu().tolist())
            self._taps_key = key
        return self._taps

    def log_params(self):
        alpha = self.act.alpha.data
        beta = alpha if self.act.__class__.__name__ == "Snake" else self.act.beta.data
        if not self.act.alpha_logscale:  # exp is baked into the kernel (cuda/activation1d.py:68-72)
            alpha, beta = torch.log(alpha), torch.log(beta)
        return alpha, beta

    def forward(self, x):
        alpha, beta = self.log_params()
        up, down = self._host_taps()
        fast = self.fast_sin if self.fast_sin is not None else (x.dtype in (torch.bfloat16, torch.float16))
        return ops.act1d(x, alpha, beta, up, down, bool(fast))
