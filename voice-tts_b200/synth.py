"""Deterministic synthetic weights and mels for parity tests and benchmarks.

There is no network for checkpoints, so every test/bench uses random-init
weights of the reference architecture.  Tensors are generated *per state-dict
key* from a generator seeded with (seed, crc32(key)), so the values do not
depend on module construction order and are identical in this container (where
they are loaded into the real reference to make golden vectors) and on the GPU
box (where only the oracle and the CUDA path exist).

State-dict key names and shapes follow the reference after
`remove_weight_norm()` (indextts/s2mel/modules/bigvgan/bigvgan.py:285-350,
388-400).
"""
import math
import zlib

import torch

from .config import in_channels, stage_channels


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def _uniform(shape, bound, seed, key):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def _normal(shape, std, seed, key):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std


def kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12):
    """Kaiser-windowed sinc low-pass taps, shape [1,1,kernel_size], fp32.

    Same arithmetic (fp32 torch ops, same order) as the reference's
    `kaiser_sinc_filter1d` (alias_free_activation/torch/filter.py:30-62) so the
    buffers a fresh `Activation1d` registers are bit-identical to the
    reference's.  Formula: Kaiser window with beta from the stop-band
    attenuation A = 2.285*(N/2-1)*pi*4*half_width + 7.95, times
    2*cutoff*sinc(2*cutoff*t), normalised to unit DC gain.
    """
    half = kernel_size // 2
    att = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    if att > 50.0:
        beta = 0.1102 * (att - 8.7)
    elif att >= 21.0:
        beta = 0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0)
    else:
        beta = 0.0
    win = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        t = torch.arange(-half, half) + 0.5
    else:
        t = torch.arange(kernel_size) - half
    if cutoff == 0:
        return torch.zeros(1, 1, kernel_size)
    taps = 2 * cutoff * win * torch.sinc(2 * cutoff * t)
    taps = taps / taps.sum()
    return taps.view(1, 1, kernel_size)


def state_dict_spec(h):
    """[(key, shape, kind)] for every tensor of the folded (no weight-norm)
    generator state dict, in the reference's naming."""
    spec = []
    c0 = h["upsample_initial_channel"]
    cin0 = in_channels(h)
    spec.append(("conv_pre.weight", (c0, cin0, 7), "conv"))
    spec.append(("conv_pre.bias", (c0,), "bias:%d" % (cin0 * 7)))
    cin = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cout = cin // 2
        # ConvTranspose1d weight is [Cin, Cout, k]; torch's fan_in for it is Cout*k
        spec.append(("ups.%d.0.weight" % i, (cin, cout, k), "conv"))
        spec.append(("ups.%d.0.bias" % i, (cout,), "bias:%d" % (cout * k)))
        cin = cout
    chans = stage_channels(h)
    n = 0
    for i, c in enumerate(chans):
        for k, dil in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            for grp in ("convs1", "convs2"):
                for j in range(len(dil)):
                    spec.append(("resblocks.%d.%s.%d.weight" % (n, grp, j), (c, c, k), "conv"))
                    spec.append(("resblocks.%d.%s.%d.bias" % (n, grp, j), (c,), "bias:%d" % (c * k)))
            for l in range(2 * len(dil)):
                p = "resblocks.%d.activations.%d" % (n, l)
                spec.append((p + ".act.alpha", (c,), "snake"))
                if h["activation"] == "snakebeta":
                    spec.append((p + ".act.beta", (c,), "snake"))
                spec.append((p + ".upsample.filter", (1, 1, 12), "filter"))
                spec.append((p + ".downsample.lowpass.filter", (1, 1, 12), "filter"))
            n += 1
    spec.append(("activation_post.act.alpha", (chans[-1],), "snake"))
    if h["activation"] == "snakebeta":
        spec.append(("activation_post.act.beta", (chans[-1],), "snake"))
    spec.append(("activation_post.upsample.filter", (1, 1, 12), "filter"))
    spec.append(("activation_post.downsample.lowpass.filter", (1, 1, 12), "filter"))
    spec.append(("conv_post.weight", (1, chans[-1], 7), "conv"))
    if h.get("use_bias_at_final", True):
        spec.append(("conv_post.bias", (1,), "bias:%d" % (chans[-1] * 7)))
    E = h.get("speaker_embedding_dim", 0)
    if E:   # speaker-conditioned v1 generator (indextts/BigVGAN/models.py:204-209): 1x1 convs on the embedding
        spec.append(("cond_layer.weight", (c0, E, 1), "conv"))
        spec.append(("cond_layer.bias", (c0,), "bias:%d" % E))
        if h.get("cond_d_vector_in_each_upsampling_layer", False):
            for i, c in enumerate(chans):
                spec.append(("conds.%d.weight" % i, (c, E, 1), "conv"))
                spec.append(("conds.%d.bias" % i, (c,), "bias:%d" % E))
    return spec


def make_state_dict(h, seed=1234, snake_std=0.5):
    """Random-init weights with PyTorch-default-like scale
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) and alpha,beta ~ N(0, snake_std) in log
    scale, so the per-channel SnakeBeta path is exercised."""
    sd = {}
    taps = kaiser_sinc_filter1d(0.25, 0.3, 12)
    for key, shape, kind in state_dict_spec(h):
        if kind == "conv":
            fan_in = shape[1] * shape[2]
            sd[key] = _uniform(shape, 1.0 / math.sqrt(fan_in), seed, key)
        elif kind.startswith("bias:"):
            sd[key] = _uniform(shape, 1.0 / math.sqrt(int(kind[5:])), seed, key)
        elif kind == "snake":
            v = _normal(shape, snake_std, seed, key)
            if not h.get("snake_logscale", True):
                v = torch.exp(v)
            sd[key] = v
        elif kind == "filter":
            sd[key] = taps.clone()
    return sd


def make_mel(batch, num_mels, frames, first_utterance=0):
    """Synthetic natural-log mels: clamp(N(-4, 2^2), -11.5, 2), seed 1000+utterance."""
    out = torch.empty(batch, num_mels, frames, dtype=torch.float32)
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(1000 + first_utterance + b)
        out[b] = (torch.randn(num_mels, frames, generator=g) * 2.0 - 4.0).clamp_(-11.5, 2.0)
    return out


def make_latent(batch, frames, gpt_dim, first_utterance=0):
    """Synthetic GPT latents for the v1 generator: N(0, 1) [B, T, gpt_dim], seed 2000+utterance."""
    out = torch.empty(batch, frames, gpt_dim, dtype=torch.float32)
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(2000 + first_utterance + b)
        out[b] = torch.randn(frames, gpt_dim, generator=g)
    return out


def make_speaker_embedding(batch, dim, first_utterance=0):
    """Synthetic speaker embeddings [B, dim] (stand-in for the ECAPA-TDNN output, models.py:213), seed 3000+utterance."""
    out = torch.empty(batch, dim, dtype=torch.float32)
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(3000 + first_utterance + b)
        out[b] = torch.randn(dim, generator=g)
    return out


# ---- s2mel tail -------------------------------------------------------------------
def s2mel_tail_spec(cfg):
    """[(key, shape, fan_in)] of the FOLDED tail weights, keys relative to the reference's DiT module with the
    `.conv.conv` nesting of encodec.SConv1d dropped (diffusion_transformer.py:139-157, wavenet.py:119-138)."""
    H, D, L, k = cfg["hidden"], cfg["dit_hidden"], cfg["n_layers"], cfg["kernel_size"]
    F, C = cfg["freq_dim"], cfg["out_channels"]
    spec = [("conv1.weight", (H, D), D), ("conv1.bias", (H,), D),
            ("t_embedder2.mlp.0.weight", (H, F), F), ("t_embedder2.mlp.0.bias", (H,), F),
            ("t_embedder2.mlp.2.weight", (H, H), H), ("t_embedder2.mlp.2.bias", (H,), H),
            ("wavenet.cond_layer.weight", (2 * H * L, H, 1), H), ("wavenet.cond_layer.bias", (2 * H * L,), H)]
    for i in range(L):
        spec.append(("wavenet.in_layers.%d.weight" % i, (2 * H, H, k), H * k))
        spec.append(("wavenet.in_layers.%d.bias" % i, (2 * H,), H * k))
        co = 2 * H if i < L - 1 else H
        spec.append(("wavenet.res_skip_layers.%d.weight" % i, (co, H, 1), H))
        spec.append(("wavenet.res_skip_layers.%d.bias" % i, (co,), H))
    spec += [("res_projection.weight", (H, D), D), ("res_projection.bias", (H,), D),
             ("final_layer.adaLN_modulation.1.weight", (2 * H, H), H), ("final_layer.adaLN_modulation.1.bias", (2 * H,), H),
             ("final_layer.linear.weight", (H, H), H), ("final_layer.linear.bias", (H,), H),
             ("conv2.weight", (C, H, 1), H), ("conv2.bias", (C,), H)]
    return spec


def make_s2mel_tail_state_dict(cfg, seed=4321):
    """Random-init folded weights, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's defaults, plus the `freqs` buffer of
    TimestepEmbedder (diffusion_transformer.py:33-37)."""
    sd = {key: _uniform(shape, 1.0 / math.sqrt(fan_in), seed, key) for key, shape, fan_in in s2mel_tail_spec(cfg)}
    half = cfg["freq_dim"] // 2
    sd["t_embedder2.freqs"] = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    return sd


def make_s2mel_tail_inputs(cfg, B, T, seed=77, lens=None):
    """x_res ~ N(0, 1) [B, T, dit_hidden], t in (0, 1) [B], t1 ~ N(0, 1) [B, hidden], x_lens [B] int32."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x_res = torch.randn(B, T, cfg["dit_hidden"], generator=g)
    t = torch.rand(B, generator=g)
    t1 = torch.randn(B, cfg["hidden"], generator=g)
    x_lens = torch.tensor(lens if lens is not None else [T] * B, dtype=torch.int32)
    return x_res, t, t1, x_lens
