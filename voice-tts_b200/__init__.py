"""B200-native BigVGAN v2 vocoder path (drop-in for
indextts/s2mel/modules/bigvgan of caishiqing/voice-tts)."""
from .config import AttrDict, load_hparams_from_json, default_hparams, tiny_hparams, v1_hparams, tiny_v1_hparams  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import voice_tts_b200.config` stays cheap
    if name in ("BigVGAN", "AMPBlock1"):
        from . import bigvgan
        return getattr(bigvgan, name)
    if name == "BigVGANv1":   # indextts/BigVGAN/models.py::BigVGAN (speaker-conditioned v1 generator)
        from . import bigvgan_v1
        return bigvgan_v1.BigVGAN
    if name in ("Activation1d", "Snake", "SnakeBeta", "UpSample1d", "DownSample1d", "LowPassFilter1d"):
        from . import activation1d
        return getattr(activation1d, name)
    raise AttributeError(name)
