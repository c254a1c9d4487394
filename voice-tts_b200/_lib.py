"""ctypes binding of libbvg_b200.so (include/bvg_b200.h).  No fallback: if the
library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("BVG_LIB_NAME", "libbvg_b200.so"))   # BVG_LIB_NAME: debug builds only

F32, BF16, F16 = 0, 1, 2
MODE_FP32, MODE_BF16 = 0, 1
SNAKE, SNAKEBETA = 0, 1
ACT_FAST_SIN = 1
UNIT_LAYERWISE, UNIT_REQUIRE_FUSED = 1, 2

_ERR = {0: "BVG_OK", -1: "BVG_EINVAL", -2: "BVG_EDTYPE", -3: "BVG_EALIGN", -4: "BVG_ECUDA",
        -5: "BVG_ENODEV", -6: "BVG_ENOMEM", -7: "BVG_ESTATE"}


class BvgConfig(ctypes.Structure):
    _fields_ = [
        ("num_mels", ctypes.c_int), ("upsample_initial_channel", ctypes.c_int),
        ("num_upsamples", ctypes.c_int), ("upsample_rates", ctypes.c_int * 8),
        ("upsample_kernel_sizes", ctypes.c_int * 8), ("num_kernels", ctypes.c_int),
        ("resblock_kernel_sizes", ctypes.c_int * 4), ("num_dilations", ctypes.c_int),
        ("resblock_dilations", (ctypes.c_int * 4) * 4), ("snake_kind", ctypes.c_int),
        ("snake_logscale", ctypes.c_int), ("use_tanh_at_final", ctypes.c_int),
        ("use_bias_at_final", ctypes.c_int), ("mode", ctypes.c_int), ("device", ctypes.c_int),
        ("input_channels_last", ctypes.c_int), ("cond_dim", ctypes.c_int), ("cond_each_up", ctypes.c_int),
    ]


class S2MelConfig(ctypes.Structure):
    _fields_ = [("hidden", ctypes.c_int), ("dit_hidden", ctypes.c_int), ("n_layers", ctypes.c_int),
                ("kernel_size", ctypes.c_int), ("dilation_rate", ctypes.c_int), ("out_channels", ctypes.c_int),
                ("freq_dim", ctypes.c_int), ("mode", ctypes.c_int), ("device", ctypes.c_int)]


_lib = None

# name -> (restype, argtypes); every symbol include/bvg_b200.h declares
_vp, _i, _i64, _fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_float)
_f = ctypes.c_float
SYMBOLS = {
    "bvg_abi_version": (_i, []),
    "bvg_last_error": (ctypes.c_char_p, []),
    "bvg_launch_count": (ctypes.c_uint64, []),
    "bvg_act1d_fwd": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _i, _i, _i64, _i, _i, _vp]),
    "bvg_act1d_cl_fwd": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _i, _i64, _i, _i, _i, _i, _vp]),
    "bvg_conv1d_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _i, _vp]),
    "bvg_conv1d_res_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i64, _i, _i, _i, _vp]),
    "bvg_conv1d_act_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _fp, _fp, _i, _i, _i, _i64, _i, _i, _i, _vp]),
    "bvg_conv1d_res_act_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _fp, _fp, _i, _i, _i, _i64, _i, _i, _i, _vp]),
    "bvg_amp_unit_fwd": (_i, [_vp] * 10 + [_fp, _fp, _vp, _f, _i, _i, _i, _i64, _i, _i, _i, _i, _vp]),
    "bvg_convtr1d_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _i, _vp]),
    "bvg_create": (_i, [ctypes.POINTER(BvgConfig), ctypes.POINTER(_vp)]),
    "bvg_destroy": (None, [_vp]),
    "bvg_set_tensor": (_i, [_vp, ctypes.c_char_p, _vp, _i64, _i]),
    "bvg_finalize": (_i, [_vp]),
    "bvg_workspace_bytes": (_i64, [_vp, _i, _i]),
    "bvg_vocoder_fwd": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "bvg_vocoder_fwd_cond": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "bvg_vocoder_fwd_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "bvg_set_option": (_i, [_vp, ctypes.c_char_p, _i]),
    "bvg_last_forward_launches": (_i, [_vp]),
    "bvg_profile_dump": (_i, [_vp, ctypes.c_char_p]),
    "bvg_profile_read": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                              ctypes.POINTER(ctypes.c_int)]),
    "bvg_s2mel_tail_create": (_i, [ctypes.POINTER(S2MelConfig), ctypes.POINTER(_vp)]),
    "bvg_s2mel_tail_destroy": (None, [_vp]),
    "bvg_s2mel_tail_set_tensor": (_i, [_vp, ctypes.c_char_p, _vp, _i64, _i]),
    "bvg_s2mel_tail_finalize": (_i, [_vp]),
    "bvg_s2mel_tail_workspace_bytes": (_i64, [_vp, _i, _i]),
    "bvg_s2mel_tail_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "bvg_s2mel_tail_set_option": (_i, [_vp, ctypes.c_char_p, _i]),
    "bvg_s2mel_tail_last_forward_launches": (_i, [_vp]),
    "bvg_cfm_euler_step": (_i, [_vp, _vp, _f, ctypes.c_double, _i, _i, _i64, _i64, _vp]),
}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libbvg_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python voice-tts_b200/build.py`; there is no CPU/PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        if not hasattr(lib, name) and os.environ.get("BVG_LIB_NAME"):   # debug builds of older revisions
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.bvg_abi_version() != 2:
        raise RuntimeError("libbvg_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().bvg_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed: %s (%s)" % (what, _ERR.get(rc, rc), msg))


def launch_count():
    return int(load().bvg_launch_count())


def taps_array(vals):
    vals = [float(v) for v in vals]
    if len(vals) != 12:
        raise RuntimeError("only 12-tap anti-aliasing filters are supported by the fused kernel (got %d)" % len(vals))
    return (ctypes.c_float * 12)(*vals)
