"""B200 path of the s2mel tail that feeds the vocoder.

`S2MelTail` holds exactly the sub-modules `DiT.forward` runs AFTER its transformer
(indextts/s2mel/modules/diffusion_transformer.py:245-256) under the reference's own state-dict names - `conv1`,
`t_embedder2`, `wavenet` (WN of wavenet.py:103-174 built from encodec.SConv1d, hence the `.conv.conv.` nesting and the
`weight_g` / `weight_v` pairs), `res_projection`, `final_layer`, `conv2` - so `tail.load_state_dict(tail_keys(dit_sd))`
takes a DiT checkpoint unchanged.  `forward` is ONE C-ABI call (`bvg_s2mel_tail_fwd`).

`solve_euler` / `CFMSolver.inference` mirror BASECFM.solve_euler / inference (flow_matching.py:31-113): the estimator
stays the caller's module (the DiT transformer is outside SURVEY section 8); the classifier-free-guidance combine, the
Euler update and the prompt zeroing of each step are one kernel (`bvg_cfm_euler_step`).  CUDA (sm_100a) only.
"""
import ctypes

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm

from . import _lib

TAIL_PREFIXES = ("conv1.", "t_embedder2.", "wavenet.", "res_projection.", "final_layer.", "conv2.")


def tail_keys(dit_state_dict):
    """the entries of a DiT state dict that belong to the tail"""
    return {k: v for k, v in dit_state_dict.items() if k.startswith(TAIL_PREFIXES)}


class _NormConv1d(nn.Module):      # encodec.NormConv1d (encodec.py:124-138) with norm='weight_norm'
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = weight_norm(nn.Conv1d(cin, cout, k))


class _SConv1d(nn.Module):         # encodec.SConv1d (encodec.py:192-228): parameter container, reflect padding is native
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = _NormConv1d(cin, cout, k)


class _WN(nn.Module):              # wavenet.py:103-138
    def __init__(self, hidden, kernel_size, n_layers):
        super().__init__()
        self.cond_layer = _SConv1d(hidden, 2 * hidden * n_layers, 1)
        self.in_layers = nn.ModuleList([_SConv1d(hidden, 2 * hidden, kernel_size) for _ in range(n_layers)])
        self.res_skip_layers = nn.ModuleList(
            [_SConv1d(hidden, 2 * hidden if i < n_layers - 1 else hidden, 1) for i in range(n_layers)])


class _TimestepEmbedder(nn.Module):  # diffusion_transformer.py:20-57
    def __init__(self, hidden, freq_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(freq_dim, hidden, bias=True), nn.SiLU(), nn.Linear(hidden, hidden, bias=True))
        import math
        half = freq_dim // 2
        self.register_buffer("freqs", torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half))


class _FinalLayer(nn.Module):      # diffusion_transformer.py:82-99 (norm_final has no parameters)
    def __init__(self, hidden):
        super().__init__()
        self.linear = weight_norm(nn.Linear(hidden, hidden, bias=True))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 2 * hidden, bias=True))


class S2MelTail(nn.Module):
    def __init__(self, cfg, precision="bf16"):
        super().__init__()
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.cfg = dict(cfg)
        self.precision = precision
        H, D = cfg["hidden"], cfg["dit_hidden"]
        self.t_embedder2 = _TimestepEmbedder(H, cfg["freq_dim"])
        self.conv1 = nn.Linear(D, H)
        self.conv2 = nn.Conv1d(H, cfg["out_channels"], 1)
        self.wavenet = _WN(H, cfg["kernel_size"], cfg["n_layers"])
        self.final_layer = _FinalLayer(H)
        self.res_projection = nn.Linear(D, H)
        self._handle = None
        self._device = None
        self._options = {}

    @classmethod
    def from_dit(cls, dit, precision="bf16"):
        """Build the tail of a reference `DiT` instance (diffusion_transformer.py:101-175 with final_layer_type 'wavenet'):
        configuration read off its sub-modules, weights loaded under its own state-dict keys, same device."""
        w = dit.wavenet
        k = w.kernel_size[0] if isinstance(w.kernel_size, (tuple, list)) else w.kernel_size   # wavenet.py:108 stores a 1-tuple
        cfg = dict(hidden=w.hidden_channels, dit_hidden=dit.conv1.in_features, n_layers=w.n_layers, kernel_size=int(k),
                   dilation_rate=w.dilation_rate, out_channels=dit.conv2.out_channels,
                   freq_dim=dit.t_embedder2.frequency_embedding_size)
        m = cls(cfg, precision=precision)
        m.load_state_dict(tail_keys(dit.state_dict()), strict=True)
        return m.to(dit.conv1.weight.device)

    def set_option(self, key, value):
        """native options ("graph": 0 / 1 / 2, "conv_own_sm"); kept across rebuilds of the native handle"""
        self._options[key] = int(value)
        if self._handle is not None:
            with torch.cuda.device(self._device):
                _lib.check(_lib.load().bvg_s2mel_tail_set_option(self._handle, key.encode(), int(value)), "bvg_s2mel_tail_set_option")

    def last_forward_launches(self):
        return int(_lib.load().bvg_s2mel_tail_last_forward_launches(self._handle)) if self._handle is not None else 0

    # ---- weights ----------------------------------------------------------------------------------------------
    def folded_state_dict(self):
        """native tensor names: weight norm folded, the `.conv.conv` nesting of SConv1d dropped"""
        sd, full = {}, self.state_dict()
        for k, v in full.items():
            if k.endswith("weight_g"):
                continue
            if k.endswith("weight_v"):
                k, v = k[:-2], torch._weight_norm(v, full[k[:-2] + "_g"], 0)
            sd[k.replace(".conv.conv.", ".")] = v
        return sd

    def load_folded_state_dict(self, sd):
        """inverse of `folded_state_dict` (weight_v = w, weight_g = ||w|| per output row): synthetic weights in tests"""
        full = {}
        for k, v in sd.items():
            kk = k
            for p in ("wavenet.cond_layer.", "wavenet.in_layers.", "wavenet.res_skip_layers."):
                if k.startswith(p):
                    head, leaf = k.rsplit(".", 1)
                    kk = head + ".conv.conv." + leaf
            if kk.endswith(".weight") and (kk.startswith("wavenet.") or kk.startswith("final_layer.linear.")):
                full[kk + "_v"] = v
                full[kk + "_g"] = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
            else:
                full[kk] = v
        return self.load_state_dict(full, strict=True)

    def _invalidate(self):
        if self._handle is not None:
            with torch.cuda.device(self._device):
                _lib.load().bvg_s2mel_tail_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def _build_native(self, device):
        lib = _lib.load()
        c = _lib.S2MelConfig()
        for f in ("hidden", "dit_hidden", "n_layers", "kernel_size", "dilation_rate", "out_channels", "freq_dim"):
            setattr(c, f, int(self.cfg[f]))
        c.mode = _lib.MODE_BF16 if self.precision == "bf16" else _lib.MODE_FP32
        c.device = device.index if device.index is not None else torch.cuda.current_device()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.bvg_s2mel_tail_create(ctypes.byref(c), ctypes.byref(handle)), "bvg_s2mel_tail_create")
        try:
            with torch.no_grad(), torch.cuda.device(device):
                tensors = [(n, t.detach().to(device=device, dtype=torch.float32).contiguous())
                           for n, t in self.folded_state_dict().items()]
                torch.cuda.current_stream(device).synchronize()   # the library packs on the legacy default stream
                for n, t in tensors:
                    _lib.check(lib.bvg_s2mel_tail_set_tensor(handle, n.encode(), t.data_ptr(), t.numel(), 1),
                               "bvg_s2mel_tail_set_tensor(%s)" % n)
            _lib.check(lib.bvg_s2mel_tail_finalize(handle), "bvg_s2mel_tail_finalize")
            for k, v in self._options.items():
                _lib.check(lib.bvg_s2mel_tail_set_option(handle, k.encode(), v), "bvg_s2mel_tail_set_option")
        except Exception:
            lib.bvg_s2mel_tail_destroy(handle)
            raise
        self._handle, self._device = handle, device

    # ---- reference semantics: diffusion_transformer.py:245-256 ---------------------------------------------------
    def forward(self, x_res, x_lens, t, t1):
        """x_res [B, T, dit_hidden] (transformer output after skip_linear), x_lens [B] or None, t [B], t1 [B, hidden]
        = DiT.t_embedder(t)  ->  [B, out_channels, T]"""
        if not x_res.is_cuda:
            raise RuntimeError("S2MelTail (B200 build) runs on CUDA only; got a %s tensor" % x_res.device)
        dev = x_res.device
        if self._handle is None or self._device != dev:
            self._invalidate()
            self._build_native(dev)
        B, T, D = x_res.shape
        if D != self.cfg["dit_hidden"] or t.shape != (B,) or t1.shape != (B, self.cfg["hidden"]):
            raise RuntimeError("S2MelTail.forward: x_res %s / t %s / t1 %s do not match the configuration" %
                               (tuple(x_res.shape), tuple(t.shape), tuple(t1.shape)))
        x_res = x_res.float().contiguous()
        t = t.to(device=dev, dtype=torch.float32).contiguous()
        t1 = t1.to(device=dev, dtype=torch.float32).contiguous()
        lens = None
        if x_lens is not None:
            if x_lens.numel() != B:
                raise RuntimeError("x_lens must have one entry per utterance")
            lens = x_lens.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty(B, self.cfg["out_channels"], T, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = _lib.load().bvg_s2mel_tail_fwd(self._handle, x_res.data_ptr(), lens.data_ptr() if lens is not None else None,
                                                t.data_ptr(), t1.data_ptr(), out.data_ptr(), B, T,
                                                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "bvg_s2mel_tail_fwd")
        return out


def euler_step_(x, dphi, dt, cfg_rate, prompt_len):
    """in place: x = x + dt * ((1 + r) * dphi[:B] - r * dphi[B:]) (r > 0; else x + dt * dphi); x[..., :prompt_len] = 0"""
    if not (x.is_cuda and dphi.is_cuda and x.dtype == torch.float32 and dphi.dtype == torch.float32):
        raise RuntimeError("euler_step_: fp32 CUDA tensors expected")
    if not (x.is_contiguous() and dphi.is_contiguous()):
        raise RuntimeError("euler_step_: contiguous tensors expected")
    B, C, T = x.shape
    if tuple(dphi.shape) != ((2 * B if cfg_rate > 0 else B), C, T):
        raise RuntimeError("euler_step_: dphi %s does not match x %s" % (tuple(dphi.shape), tuple(x.shape)))
    with torch.cuda.device(x.device):
        rc = _lib.load().bvg_cfm_euler_step(x.data_ptr(), dphi.data_ptr(), float(dt), float(cfg_rate), B, C, T, int(prompt_len),
                                            torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "bvg_cfm_euler_step")
    return x


@torch.no_grad()
def solve_euler(estimator, x, x_lens, prompt, mu, style, f0, t_span, inference_cfg_rate=0.5, zero_prompt_speech_token=False):
    """BASECFM.solve_euler (flow_matching.py:57-113) with the same arguments; `estimator(x, prompt_x, x_lens, t, style, mu)`
    is the caller's DiT.  The time bookkeeping (`t`, `dt`) is done on fp32 CPU scalars exactly as the reference does on
    0-dim tensors, so the trajectory is bit-identical for a bit-identical estimator."""
    t_span = t_span.detach().to("cpu", torch.float32)
    t = t_span[0]
    prompt_len = prompt.size(-1)
    prompt_x = torch.zeros_like(x)
    prompt_x[..., :prompt_len] = prompt[..., :prompt_len]
    x = x.float().contiguous().clone()
    x[..., :prompt_len] = 0
    if zero_prompt_speech_token:
        mu[..., :prompt_len] = 0
    for step in range(1, len(t_span)):
        dt = t_span[step] - t_span[step - 1]      # (the reference's `dt = t_span[step + 1] - t` at the end of an iteration
        tt = t.to(x.device).unsqueeze(0)          #  is overwritten here before it is used, flow_matching.py:86,110-111)
        if inference_cfg_rate > 0:
            dphi = estimator(torch.cat([x, x], dim=0), torch.cat([prompt_x, torch.zeros_like(prompt_x)], dim=0), x_lens,
                             torch.cat([tt, tt], dim=0), torch.cat([style, torch.zeros_like(style)], dim=0),
                             torch.cat([mu, torch.zeros_like(mu)], dim=0))
        else:
            dphi = estimator(x, prompt_x, x_lens, tt, style, mu)
        euler_step_(x, dphi.float().contiguous(), float(dt), inference_cfg_rate, prompt_len)
        t = t + dt
    return x


class CFMSolver:
    """`inference(mu, x_lens, prompt, style, f0, n_timesteps, temperature, inference_cfg_rate)` of BASECFM
    (flow_matching.py:31-55) around an injected estimator."""

    def __init__(self, estimator, in_channels=80, zero_prompt_speech_token=False):
        self.estimator, self.in_channels, self.zero_prompt_speech_token = estimator, in_channels, zero_prompt_speech_token

    @torch.no_grad()
    def inference(self, mu, x_lens, prompt, style, f0, n_timesteps, temperature=1.0, inference_cfg_rate=0.5):
        B, T = mu.size(0), mu.size(1)
        z = torch.randn([B, self.in_channels, T], device=mu.device) * temperature
        t_span = torch.linspace(0, 1, n_timesteps + 1, device=mu.device)
        return solve_euler(self.estimator, z, x_lens, prompt, mu, style, f0, t_span, inference_cfg_rate,
                           self.zero_prompt_speech_token)
