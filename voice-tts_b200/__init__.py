"""B200-native BigVGAN v2 vocoder path (drop-in for
indextts/s2mel/modules/bigvgan of caishiqing/voice-tts)."""
from .config import AttrDict, load_hparams_from_json, default_hparams, tiny_hparams  # noqa: F401
