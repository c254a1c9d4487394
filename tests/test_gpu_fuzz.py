"""Randomised generator configurations against the oracle (fixed seeds, so the cases are reproducible): channel counts that
are not multiples of 16 / 32 / 64, every odd resblock kernel 3..11 with dilations 1..7, even and odd upsampling rates with
kernel u, 2u and other 3-tap polyphase shapes, both snake kinds and scales, clamp / tanh, with and without final bias, ragged
batch / length combinations down to T = 1.  fp32 mode <= 1e-5 of max-abs (every case); bf16 mode by SNR (narrow random-weight
generators: 30 dB); bf16x3 mode >= 70 dB."""
import os
import random

import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def random_case(seed, cfg):
    r = random.Random(seed)
    nst = r.choice([1, 2, 3])
    rates, ksz = [], []
    for _ in range(nst):
        u = r.choice([2, 2, 3, 4, 4, 5, 8])
        pads = [p for p in range(0, u + 1) if 2 * p + u <= 2 * u + p]          # k = u + 2p, pad p <= u
        p = r.choice(pads)
        rates.append(u)
        ksz.append(u + 2 * p)
    c_last = r.choice([8, 12, 16, 20, 24, 40, 48, 56, 72, 96])
    c0 = c_last * (2 ** nst)
    nk = r.choice([1, 2, 3])
    nd = r.choice([1, 2, 3])
    rk = [r.choice([3, 5, 7, 9, 11]) for _ in range(nk)]
    rd = [[r.choice([1, 2, 3, 5, 7]) for _ in range(nd)] for _ in range(nk)]
    # keep the conv halo within the native plan's budget ((k-1)*d <= 64)
    rd = [[d if (k - 1) * d <= 64 else 1 for d in ds] for k, ds in zip(rk, rd)]
    h = cfg.default_hparams(upsample_initial_channel=c0, num_mels=r.choice([5, 16, 33, 80]), upsample_rates=rates,
                            upsample_kernel_sizes=ksz, resblock_kernel_sizes=rk, resblock_dilation_sizes=rd,
                            activation=r.choice(["snakebeta", "snakebeta", "snake"]), snake_logscale=r.choice([True, True, False]),
                            use_tanh_at_final=r.choice([True, False]), use_bias_at_final=r.choice([True, False]))
    B = r.choice([1, 1, 2, 3])
    T = r.choice([1, 2, 3, 7, 19, 40, 77])
    return h, B, T


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("BVG_FUZZ_SEEDS", "20")))))   # BVG_FUZZ_SEEDS=200 for a longer hunt
def test_random_generator_vs_oracle(pkg, synth, cfg, seed):
    h, B, T = random_case(seed, cfg)
    sd = synth.make_state_dict(h, seed=100 + seed)
    mel = synth.make_mel(B, h["num_mels"], T)
    ref = O.generator_forward(sd, h, mel)
    scale = float(ref.abs().max())
    for precision in ("fp32", "bf16", "bf16x3"):
        m = pkg.BigVGAN(h, precision=precision)
        m.remove_weight_norm()
        m.load_state_dict(sd)
        m = m.to(DEV).eval()
        with torch.no_grad():
            wav = m(mel.to(DEV)).cpu()
        assert wav.shape == ref.shape, (seed, dict(h))
        assert torch.isfinite(wav).all()
        if precision == "fp32":
            assert (wav - ref).abs().max() <= 1e-5 * scale, (seed, dict(h), B, T)
        else:
            snr = O.snr_db(ref, wav)
            assert snr >= (30.0 if precision == "bf16" else 70.0), (seed, precision, snr, dict(h), B, T)
