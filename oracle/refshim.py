"""ORACLE support — imports the *real* reference vocoder (read-only tree at
/root/reference) in the build container.  Test infrastructure only; never
imported by the product.  `/root/reference` does not exist on the GPU box, so
everything here is gated on `available()`.

The reference's bigvgan package imports matplotlib and librosa at module import
(indextts/s2mel/modules/bigvgan/utils.py:6,11, meldataset.py:13,15) although the
generator never calls them; they are absent in this image, so empty stand-in
modules are registered before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BVG_REFERENCE_ROOT", "/root/reference")
_CONFIG = "indextts/s2mel/modules/bigvgan/config.json"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _CONFIG))


def load():
    """Returns the reference's `bigvgan` module (classes BigVGAN, AMPBlock1 ...)."""
    for name in ("matplotlib", "matplotlib.pylab", "librosa", "librosa.util", "librosa.filters"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    sys.modules["librosa.util"].normalize = None
    sys.modules["librosa.filters"].mel = None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)
    from indextts.s2mel.modules.bigvgan import bigvgan  # noqa: E402
    return bigvgan


def build_generator(h, state_dict):
    """Reference BigVGAN with `state_dict` (folded keys) loaded, eval mode."""
    import contextlib
    import io
    mod = load()
    hh = mod.AttrDict(dict(h))
    with contextlib.redirect_stdout(io.StringIO()):
        m = mod.BigVGAN(hh, use_cuda_kernel=False)
        m.remove_weight_norm()
    missing = m.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()
