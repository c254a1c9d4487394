"""Hyper-parameter container for the BigVGAN v2 vocoder path.

Mirrors the reference's `AttrDict` / `load_hparams_from_json`
(indextts/s2mel/modules/bigvgan/env.py:8-11, bigvgan.py:25-28) so that a
`config.json` written by the reference's `_save_pretrained` loads unchanged.
"""
import json

# Keys of indextts/s2mel/modules/bigvgan/config.json that the generator reads
# (bigvgan.py:266-358).  Training-only keys are accepted and ignored.
BIGVGAN_V2_22KHZ_80BAND_256X = {
    "resblock": "1",
    "upsample_rates": [4, 4, 2, 2, 2, 2],
    "upsample_kernel_sizes": [8, 8, 4, 4, 4, 4],
    "upsample_initial_channel": 1536,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "use_tanh_at_final": False,
    "use_bias_at_final": False,
    "activation": "snakebeta",
    "snake_logscale": True,
    "num_mels": 80,
    "hop_size": 256,
    "sampling_rate": 22050,
}


# The IndexTTS-v1 speaker-conditioned generator (indextts/BigVGAN/models.py:130-250).  Its hyper-parameters live in
# `checkpoints/config.yaml` (section `bigvgan`), which the reference downloads at deploy time and does NOT ship in the
# tree; these are the values of the published IndexTTS-1.x configuration (gpt_dim 1024 for v1.0, 1280 for v1.5; x1024
# upsampling of the GPT latent to 24 kHz).  Two stages use kernel == stride (padding 0), which the v2 plan never does.
INDEXTTS_V1_BIGVGAN = {
    "resblock": "1",
    "upsample_rates": [4, 4, 4, 4, 2, 2],
    "upsample_kernel_sizes": [8, 8, 4, 4, 4, 4],
    "upsample_initial_channel": 1536,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "feat_upsample": False,
    "speaker_embedding_dim": 512,
    "cond_d_vector_in_each_upsampling_layer": True,
    "gpt_dim": 1024,
    "activation": "snakebeta",
    "snake_logscale": True,
    "use_tanh_at_final": True,      # models.py:248 `torch.tanh` unconditionally
    "use_bias_at_final": True,      # models.py:192 conv_post keeps its bias
    "num_mels": 100,                # width of the speaker encoder's reference mel, not of the generator input
    "hop_size": 256,
    "sampling_rate": 24000,
}


class AttrDict(dict):
    """dict whose keys are also attributes (h.num_mels and h["num_mels"])."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


# The s2mel tail: `s2mel.wavenet` / `s2mel.DiT` of the published IndexTTS-2 config.yaml
# (the reference tree does not ship checkpoints/config.yaml; values restated from the release - every test is parameterised).
INDEXTTS2_S2MEL_TAIL = {"hidden": 512, "dit_hidden": 512, "n_layers": 8, "kernel_size": 5, "dilation_rate": 1,
                        "out_channels": 80, "freq_dim": 256}


def s2mel_tail_config(**overrides):
    c = dict(INDEXTTS2_S2MEL_TAIL)
    c.update(overrides)
    return c


def load_hparams_from_json(path) -> AttrDict:
    with open(path) as f:
        return AttrDict(json.load(f))


def default_hparams(**overrides) -> AttrDict:
    h = AttrDict(json.loads(json.dumps(BIGVGAN_V2_22KHZ_80BAND_256X)))
    h.update(overrides)
    return h


def tiny_hparams(**overrides) -> AttrDict:
    """A shrunken generator (4 stages 96/48/24/12 channels, x64 upsampling, the
    same 3 kernel sizes x 3 dilations) used by the fast parity tests and the
    committed golden vectors.  12 channels exercises the channel-padding path."""
    h = default_hparams(upsample_initial_channel=192, num_mels=16,
                        upsample_rates=[4, 4, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4])
    h.update(overrides)
    return h


def v1_hparams(**overrides) -> AttrDict:
    h = AttrDict(json.loads(json.dumps(INDEXTTS_V1_BIGVGAN)))
    h.update(overrides)
    return h


def tiny_v1_hparams(**overrides) -> AttrDict:
    """Shrunken v1 generator: a k = 2u stage, a k = u stage and a k = 4/u = 2 stage, 3 kernel sizes x 3 dilations (the v1
    AMPBlock1 hard-codes three dilations, models.py:24-34), conditioning in every upsampling layer."""
    h = v1_hparams(upsample_initial_channel=96, gpt_dim=40, speaker_embedding_dim=24,
                   upsample_rates=[4, 4, 2], upsample_kernel_sizes=[8, 4, 4])
    h.update(overrides)
    return h


def in_channels(h):
    """input channels of conv_pre: the GPT latent width for the v1 generator (models.py:149), else num_mels."""
    return h["gpt_dim"] if h.get("gpt_dim") else h["num_mels"]


def stage_channels(h):
    c0 = h["upsample_initial_channel"]
    return [c0 // (2 ** (i + 1)) for i in range(len(h["upsample_rates"]))]


def total_upsample(h):
    n = 1
    for u in h["upsample_rates"]:
        n *= u
    return n


def macs_per_frame(h):
    """Dense-conv multiply-accumulates per mel frame."""
    c0 = h["upsample_initial_channel"]
    macs = in_channels(h) * c0 * 7
    t = 1
    cin = c0
    for u, ku in zip(h["upsample_rates"], h["upsample_kernel_sizes"]):
        cout = cin // 2
        macs += cin * cout * ku * t  # ConvTranspose1d: Cin*Cout*k per input sample
        t *= u
        for k, dil in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            macs += 2 * len(dil) * cout * cout * k * t
        cin = cout
    macs += cin * 1 * 7 * t
    return macs


def act_elems_per_frame(h):
    """Elements that pass through an Activation1d per mel frame."""
    n = 0
    t = 1
    for u, c in zip(h["upsample_rates"], stage_channels(h)):
        t *= u
        n += c * t * 2 * sum(len(d) for d in h["resblock_dilation_sizes"])
    n += stage_channels(h)[-1] * t
    return n
