"""Drop-in `BigVGAN` generator (mel -> waveform) for indextts/infer_v2.py.

Same constructor, `forward`, `remove_weight_norm`, `from_pretrained`,
`_save_pretrained` and state-dict key names as the reference
(indextts/s2mel/modules/bigvgan/bigvgan.py:243-492).  Parameters live in ordinary
torch modules so `load_state_dict` / `.to(device)` / checkpoints work unchanged;
`forward` hands the folded weights to the native library once and then runs the
whole generator as ONE C-ABI call (`bvg_vocoder_fwd`).  CUDA (sm_100a) only -
there is no torch fallback.
"""
import json
import os

import torch
import torch.nn as nn
from torch.nn import Conv1d, ConvTranspose1d
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _lib, ops
from .activation1d import Activation1d, Snake, SnakeBeta
from .config import AttrDict, in_channels, load_hparams_from_json, total_upsample


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def _make_act(h, channels):
    if h["activation"] == "snake":
        return Snake(channels, alpha_logscale=h["snake_logscale"])
    if h["activation"] == "snakebeta":
        return SnakeBeta(channels, alpha_logscale=h["snake_logscale"])
    raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")


class AMPBlock1(nn.Module):
    """Parameter container of one AMP block (bigvgan.py:31-147).  Stand-alone
    `forward` runs layer by layer through the single-layer ops; inside `BigVGAN`
    the native plan executes the block."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5), activation=None):
        super().__init__()
        self.h = h
        self.dilation = tuple(dilation)
        self.convs1 = nn.ModuleList([
            weight_norm(Conv1d(channels, channels, kernel_size, stride=1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilation])
        self.convs2 = nn.ModuleList([
            weight_norm(Conv1d(channels, channels, kernel_size, stride=1, dilation=1, padding=get_padding(kernel_size, 1)))
            for _ in dilation])
        self.num_layers = len(self.convs1) + len(self.convs2)
        hh = dict(h)
        hh["activation"] = activation or h["activation"]
        self.activations = nn.ModuleList([Activation1d(activation=_make_act(hh, channels)) for _ in range(self.num_layers)])

    def forward(self, x, precision="fp32"):
        """xt = a1(x); xt = c1(xt); xt = a2(xt); xt = c2(xt); x = xt + x per dilation (bigvgan.py:132-141): one
        `bvg_amp_unit_fwd` call per unit - ONE kernel in bf16 mode for <= 96 channels, the four layers otherwise."""
        acts1, acts2 = self.activations[::2], self.activations[1::2]
        empty = x.new_empty(0)
        for c1, c2, a1, a2, d in zip(self.convs1, self.convs2, acts1, acts2, self.dilation):
            t1, t2 = a1._host_taps(), a2._host_taps()
            if t1 == t2 and a1.fast_sin is None and a2.fast_sin is None:
                al1, be1 = a1.log_params()
                al2, be2 = a2.log_params()
                x = ops.amp_unit(x, _folded_weight(c1), c1.bias if c1.bias is not None else empty, _folded_weight(c2),
                                 c2.bias if c2.bias is not None else empty, al1, be1, al2, be2, t1[0], t1[1], empty, 1.0, False,
                                 d, precision, 0)
                continue
            xt = a1(x)      # units whose two activations carry different filters: layer by layer through the single-layer ops
            xt = ops.conv1d(xt, _folded_weight(c1), c1.bias, d, precision, 0)
            xt = a2(xt)
            xt = ops.conv1d(xt, _folded_weight(c2), c2.bias, 1, precision, 0)
            x = xt + x
        return x

    def remove_weight_norm(self):
        for l in list(self.convs1) + list(self.convs2):
            remove_weight_norm(l)


def _folded_weight(conv):
    """weight of a conv whether or not legacy weight_norm is still attached."""
    if hasattr(conv, "weight_g"):
        return torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    return conv.weight


class BigVGAN(nn.Module):
    def __init__(self, h, use_cuda_kernel: bool = False, precision: str = None):
        super().__init__()
        if not isinstance(h, AttrDict):
            h = AttrDict(dict(h))
        self.h = h
        self.h["use_cuda_kernel"] = use_cuda_kernel  # accepted for compatibility: this build is always the CUDA path
        self.precision = precision or h.get("bvg_precision") or os.environ.get("BVG_PRECISION", "bf16")
        # "bf16": bf16 operands on the tensor cores (>= 40 dB); "fp32": fp32 SIMT convolutions (<= 1e-5 of the reference);
        # "bf16x3": fp32 storage, every convolution as three (option split_terms: 6, 9) bf16 tensor-core passes over three-term
        # splits of both operands: 93-96 dB / 2e-5 on the full generator - the tensor cores' fp32 accumulation is the limit -
        # at 8x (4.5x, 3x) the fp32 mode's speed
        if self.precision not in ("bf16", "fp32", "bf16x3"):
            raise ValueError("precision must be 'bf16', 'fp32' or 'bf16x3'")
        if h["resblock"] != "1":
            # AMPBlock2.forward has no return in the reference (bigvgan.py:232-236): not a usable configuration
            raise ValueError("Incorrect resblock class specified in hyperparameters. Got %s" % h["resblock"])
        self.num_kernels = len(h["resblock_kernel_sizes"])
        self.num_upsamples = len(h["upsample_rates"])
        c0 = h["upsample_initial_channel"]
        self.conv_pre = weight_norm(Conv1d(in_channels(h), c0, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            self.ups.append(nn.ModuleList([
                weight_norm(ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2))]))
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
                self.resblocks.append(AMPBlock1(h, ch, k, d, activation=h["activation"]))
        self.activation_post = Activation1d(activation=_make_act(h, ch))
        self.use_bias_at_final = h.get("use_bias_at_final", True)
        self.conv_post = weight_norm(Conv1d(ch, 1, 7, 1, padding=3, bias=self.use_bias_at_final))
        self.use_tanh_at_final = h.get("use_tanh_at_final", True)
        self._hid = None
        self._worker_hids = []       # extra native handles of forward_segments (one per concurrent length group)
        self._worker_streams = []
        self._options = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # ---- native handle management -------------------------------------------------------------
    def _invalidate(self):
        if self._hid is not None:
            ops.release_handle(self._hid)
            self._hid = None
        for hid in getattr(self, "_worker_hids", []):
            ops.release_handle(hid)
        self._worker_hids = []

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def set_option(self, key, value):
        """native options: graph, conv_impl (0 auto / 1 simt / 2 tcgen05), fast_sin, workspace_mb"""
        self._options[key] = int(value)
        for hid in ([self._hid] if self._hid is not None else []) + list(self._worker_hids):
            _lib.check(_lib.load().bvg_set_option(ops._HANDLES[hid][0], key.encode(), int(value)), "bvg_set_option")

    def folded_state_dict(self):
        """state dict with weight norm folded (the keys `remove_weight_norm()` leaves)."""
        sd = {}
        for k, v in self.state_dict().items():
            if k.endswith(".weight_g") or k.startswith("speaker_encoder."):
                continue
            if k.endswith(".weight_v"):
                g = self.state_dict()[k[:-2] + "_g"]
                sd[k[:-9] + ".weight"] = torch._weight_norm(v, g, 0)
            else:
                sd[k] = v
        return sd

    def _build_native(self, device):
        self._hid = self._make_handle(device)

    def _make_handle(self, device):
        import ctypes
        h = self.h
        lib = _lib.load()
        cfg = _lib.BvgConfig()
        cfg.num_mels = in_channels(h)
        cfg.input_channels_last, cfg.cond_dim, cfg.cond_each_up = self._native_conditioning()
        cfg.upsample_initial_channel = h["upsample_initial_channel"]
        cfg.num_upsamples = self.num_upsamples
        if self.num_upsamples > 8 or self.num_kernels > 4:
            raise RuntimeError("configuration exceeds the native plan's limits (8 stages, 4 kernel sizes)")
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernel_sizes[i] = k
        cfg.num_kernels = self.num_kernels
        nd = len(h["resblock_dilation_sizes"][0])
        cfg.num_dilations = nd
        for j, (k, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            cfg.resblock_kernel_sizes[j] = k
            if len(dil) != nd or nd > 4:
                raise RuntimeError("all resblocks must have the same number (<=4) of dilations")
            for l, d in enumerate(dil):
                cfg.resblock_dilations[j][l] = d
        cfg.snake_kind = _lib.SNAKEBETA if h["activation"] == "snakebeta" else _lib.SNAKE
        cfg.snake_logscale = 1 if h["snake_logscale"] else 0
        cfg.use_tanh_at_final = 1 if self.use_tanh_at_final else 0
        cfg.use_bias_at_final = 1 if self.use_bias_at_final else 0
        cfg.mode = _lib.MODE_BF16 if self.precision == "bf16" else _lib.MODE_FP32
        cfg.device = device.index if device.index is not None else torch.cuda.current_device()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.bvg_create(ctypes.byref(cfg), ctypes.byref(handle)), "bvg_create")
        try:
            with torch.no_grad(), torch.cuda.device(device):
                # the folded / converted tensors are produced on torch's current stream; bvg_set_tensor packs them on the
                # legacy default stream, which does not wait for a non-blocking side stream: order the two explicitly
                tensors = [(name, t.detach().to(device=device, dtype=torch.float32).contiguous())
                           for name, t in self.folded_state_dict().items()]
                torch.cuda.current_stream(device).synchronize()
                for name, t in tensors:
                    _lib.check(lib.bvg_set_tensor(handle, name.encode(), t.data_ptr(), t.numel(), 1),
                               "bvg_set_tensor(%s)" % name)
            _lib.check(lib.bvg_finalize(handle), "bvg_finalize")
            if self.precision == "bf16x3":
                _lib.check(lib.bvg_set_option(handle, b"conv_impl", 3), "bvg_set_option")
            for k, v in self._options.items():
                _lib.check(lib.bvg_set_option(handle, k.encode(), v), "bvg_set_option")
        except Exception:
            lib.bvg_destroy(handle)
            raise
        return ops.register_handle(handle, in_channels(h), total_upsample(h), cfg.device)

    def _native_conditioning(self):
        """(input_channels_last, cond_dim, cond_each_up) of the native plan; the v1 subclass overrides it."""
        return 0, 0, 0

    # ---- reference API ----------------------------------------------------------------------------
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("BigVGAN (B200 build) runs on CUDA only; got a %s tensor" % x.device)
        if not x.is_floating_point():
            raise RuntimeError("BigVGAN expects a floating-point mel, got %s" % x.dtype)
        if self._hid is None or ops._HANDLES[self._hid][3] != x.device.index:
            self._invalidate()
            self._build_native(x.device)
        # infer_v2.py:735 passes `vc_target.float()`; the v1 pipeline calls its vocoder under fp16 autocast (infer.py:474):
        # half / bfloat16 mels are accepted and widened, the generator's arithmetic is set by `precision`
        return ops.vocoder(x.float().contiguous(), self._hid)

    def forward_host(self, mel_cpu, int16=False, out=None):
        """host mel -> host wav through `bvg_vocoder_fwd_host` (H2D + generator + D2H in one
        native call; int16 applies infer_v2.py:740's clamp(32767*wav))."""
        if mel_cpu.is_cuda or mel_cpu.dtype != torch.float32:
            raise RuntimeError("forward_host expects a float32 CPU tensor")
        dev = next(self.parameters()).device
        if not dev.type == "cuda":
            raise RuntimeError("move the model to a CUDA device first")
        if self._hid is None:
            self._build_native(dev)
        if mel_cpu.dim() != 3 or mel_cpu.shape[1] != in_channels(self.h):
            raise RuntimeError("forward_host expects mel [B, %d, T], got %s" % (in_channels(self.h), tuple(mel_cpu.shape)))
        mel_cpu = mel_cpu.contiguous()
        B, _, T = mel_cpu.shape
        n = T * total_upsample(self.h)
        want = torch.int16 if int16 else torch.float32
        if out is None:
            out = torch.empty(B, 1, n, dtype=want)
        elif out.is_cuda or out.dtype != want or not out.is_contiguous() or out.numel() != B * n:
            # the native call writes B*T*hop elements of `want` straight into this buffer
            raise RuntimeError("forward_host: `out` must be a contiguous CPU %s tensor with %d elements" % (want, B * n))
        with torch.cuda.device(dev):
            rc = _lib.load().bvg_vocoder_fwd_host(ops._HANDLES[self._hid][0], mel_cpu.data_ptr(), out.data_ptr(),
                                                  1 if int16 else 0, B, T, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "bvg_vocoder_fwd_host")
        return out

    def forward_segments(self, mels, concurrency=2):
        """Vocode a list of segments of DIFFERENT lengths in as few launches as exactness allows.

        `infer_v2.py:616-744` runs the vocoder once per text segment (`wav = self.bigvgan(vc_target.float())`).  Segments of
        equal length are stacked into one batch; every distinct length is its own batch, because the convolutions zero-pad
        and the activations replicate-pad at each utterance's TRUE end - padding a short mel to a common length would change
        its last ~34 frames.  Each returned waveform is bit-identical to `self(mel)` on that segment alone (batch rows are
        independent, tests/test_gpu_vocoder.py).  mels: tensors [num_mels, T_i] or [1, num_mels, T_i] on one CUDA device;
        returns a list of [1, T_i * hop] tensors in the input order."""
        if not mels:
            return []
        items = [m if m.dim() == 3 else m.unsqueeze(0) for m in mels]
        for m in items:
            if m.dim() != 3 or m.shape[0] != 1:
                raise RuntimeError("forward_segments expects [num_mels, T] or [1, num_mels, T] tensors")
        groups = {}
        for i, m in enumerate(items):
            groups.setdefault(int(m.shape[-1]), []).append(i)
        out = [None] * len(items)
        for i in groups.pop(0, []):
            out[i] = items[i].new_empty(1, 0)
        todo = sorted(groups.items(), key=lambda kv: -kv[0] * len(kv[1]))       # largest group first
        workers = max(1, min(int(concurrency), len(todo)))
        if workers == 1:
            for T, idx in todo:
                wav = self.forward(torch.cat([items[i] for i in idx], dim=0).float())
                for row, i in enumerate(idx):
                    out[i] = wav[row]
            return out
        # Different lengths cannot share a batch (see above) but they can share the GPU: a single 3-15 s segment leaves SMs idle
        # in the wide stages.  Groups are dealt to `workers` native handles (own workspace, weights packed once per handle),
        # each on its own stream that forks from and joins the caller's stream; per segment the kernels and their order are
        # exactly those of `self(mel)`, so the results stay bit-identical.
        dev = items[0].device
        if self._hid is None or ops._HANDLES[self._hid][3] != dev.index:
            self._invalidate()
            self._build_native(dev)
        while len(self._worker_hids) < workers - 1:
            self._worker_hids.append(self._make_handle(dev))
        if len(self._worker_streams) < workers - 1 or any(s.device != dev for s in self._worker_streams):
            self._worker_streams = [torch.cuda.Stream(device=dev) for _ in range(workers - 1)]
        load = [0] * workers
        plan = [[] for _ in range(workers)]
        for T, idx in todo:                                                      # longest-processing-time-first assignment
            w = load.index(min(load))
            plan[w].append((T, idx))
            load[w] += T * len(idx)
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        for w in range(workers):
            stream = cur if w == 0 else self._worker_streams[w - 1]
            hid = self._hid if w == 0 else self._worker_hids[w - 1]
            with torch.cuda.stream(stream):
                if w:
                    stream.wait_event(fork)
                for T, idx in plan[w]:
                    wav = ops.vocoder(torch.cat([items[i] for i in idx], dim=0).float().contiguous(), hid)
                    if w:
                        wav.record_stream(cur)
                    for row, i in enumerate(idx):
                        out[i] = wav[row]
        for w in range(1, workers):
            cur.wait_stream(self._worker_streams[w - 1])
        return out

    def read_profile(self):
        """{category: (ms, algorithmic work, launches)} since the last read (option profile=1)."""
        import ctypes
        out = {}
        if self._hid is None:
            return out
        # category 3 ("other") clears the records: read it last
        for cat, name in ((0, "conv_tcgen05"), (1, "conv_simt"), (2, "activation"), (4, "amp_unit"), (3, "other")):
            ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
            _lib.check(_lib.load().bvg_profile_read(ops._HANDLES[self._hid][0], cat, ctypes.byref(ms),
                                                    ctypes.byref(work), ctypes.byref(n)), "bvg_profile_read")
            out[name] = (ms.value, work.value, n.value)
        return out

    def dump_profile(self, path):
        _lib.check(_lib.load().bvg_profile_dump(ops._HANDLES[self._hid][0], str(path).encode()), "bvg_profile_dump")

    def last_forward_launches(self):
        return 0 if self._hid is None else int(_lib.load().bvg_last_forward_launches(ops._HANDLES[self._hid][0]))

    def remove_weight_norm(self):
        try:
            print("Removing weight norm...")
            for l in self.ups:
                for l_i in l:
                    remove_weight_norm(l_i)
            for l in self.resblocks:
                l.remove_weight_norm()
            remove_weight_norm(self.conv_pre)
            remove_weight_norm(self.conv_post)
        except ValueError:
            print("[INFO] Model already removed weight norm. Skipping!")
        self._invalidate()

    def _save_pretrained(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        torch.save({"generator": self.state_dict()}, os.path.join(save_directory, "bigvgan_generator.pt"))
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(dict(self.h), f, indent=4)

    save_pretrained = _save_pretrained

    @classmethod
    def from_pretrained(cls, model_id, use_cuda_kernel=False, map_location="cpu", precision=None, **hub_kwargs):
        """model_id: a local directory with config.json + bigvgan_generator.pt, or a hub repo id
        (needs huggingface_hub and network, like the reference's _from_pretrained)."""
        if os.path.isdir(model_id):
            config_file = os.path.join(model_id, "config.json")
            model_file = os.path.join(model_id, "bigvgan_generator.pt")
        else:
            from huggingface_hub import hf_hub_download
            config_file = hf_hub_download(repo_id=model_id, filename="config.json", **hub_kwargs)
            model_file = hf_hub_download(repo_id=model_id, filename="bigvgan_generator.pt", **hub_kwargs)
        h = load_hparams_from_json(config_file)
        model = cls(h, use_cuda_kernel=use_cuda_kernel, precision=precision)
        ckpt = torch.load(model_file, map_location=map_location)
        try:
            model.load_state_dict(ckpt["generator"])
        except RuntimeError:
            print("[INFO] the pretrained checkpoint does not contain weight norm. Loading the checkpoint after removing weight norm!")
            model.remove_weight_norm()
            model.load_state_dict(ckpt["generator"])
        return model
