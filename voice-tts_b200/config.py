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


class AttrDict(dict):
    """dict whose keys are also attributes (h.num_mels and h["num_mels"])."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


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
    macs = h["num_mels"] * c0 * 7
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
