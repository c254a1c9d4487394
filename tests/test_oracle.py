"""CPU tests: the oracle against the golden vectors produced by the real
reference (oracle/make_golden.py), against the survey's known-answer facts
(SURVEY.md 8(c)), and - when /root/reference is present - against the reference
itself."""
import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O
from oracle import refshim
from oracle.make_golden import ACT_CASES, sd_fingerprint

TAP_BITS = [0x3B04F861, 0x3C19D646, 0xBCD14084, 0xBD6C2A26, 0x3E03A888, 0x3EE2EC65]


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_filter_taps_bit_patterns(golden, synth):
    g = golden("activation1d")["taps"]
    for taps in (O.kaiser_taps().numpy(), synth.kaiser_sinc_filter1d().reshape(-1).numpy()):
        bits = taps.view(np.uint32)
        assert list(bits[:6]) == TAP_BITS
        assert list(bits[6:]) == TAP_BITS[::-1]          # symmetric f[k] = f[11-k]
        assert np.array_equal(taps, g)
    assert abs(float(g[0::2].sum()) - 0.5) < 1e-7 and abs(float(g[1::2].sum()) - 0.5) < 1e-7


@pytest.mark.parametrize("case", ACT_CASES, ids=[c[0] for c in ACT_CASES])
def test_activation1d_vs_reference_golden(golden, case):
    name, B, C, T, kind, logscale, scale = case
    g = golden("activation1d")
    x, a, b = t(g[name + ".x"]), t(g[name + ".alpha"]), t(g[name + ".beta"])
    taps = t(g["taps"])
    y64 = O.activation1d(x.double(), a.double(), b.double(), taps.double(), taps.double(), logscale)
    assert (y64 - t(g[name + ".y64"])).abs().max() < 1e-12
    y32 = O.activation1d(x, a, b, taps, taps, logscale)
    assert (y32 - t(g[name + ".y"])).abs().max() < 2e-6 * max(1.0, float(x.abs().max()))
    ys = O.activation1d_staged(x.double(), a.double(), b.double(), taps.double(), taps.double(), logscale)
    assert (ys - y64).abs().max() < 1e-12


def test_activation1d_dc_gain_and_linearity():
    # zero periodic part (beta -> +inf) makes the operator a pure up/down FIR
    # cascade: constants pass through exactly (unit DC gain incl. replicate edges)
    taps = O.kaiser_taps(dtype=torch.float64)
    x = torch.full((1, 2, 33), 0.75, dtype=torch.float64)
    big = torch.full((2,), 40.0, dtype=torch.float64)
    y = O.activation1d(x, torch.zeros(2, dtype=torch.float64), big, taps, taps)
    assert (y - 0.75).abs().max() < 1e-12
    x1, x2 = torch.randn(1, 2, 50, dtype=torch.float64), torch.randn(1, 2, 50, dtype=torch.float64)
    f = lambda z: O.activation1d(z, torch.zeros(2, dtype=torch.float64), big, taps, taps)
    assert (f(x1 + 2 * x2) - f(x1) - 2 * f(x2)).abs().max() < 1e-12


@pytest.mark.parametrize("tag", ["c16_k3", "c16_k7", "c8_k11"])
def test_ampblock1_vs_reference_golden(golden, cfg, tag):
    g = golden("ampblock1")
    sd = {k[len(tag) + 4:]: t(g[k]) for k in g.files if k.startswith(tag + ".sd.")}
    sd = {"rb." + k: v for k, v in sd.items()}
    h = cfg.default_hparams()
    y = O.amp_block1(sd, "rb", t(g[tag + ".x"]), h, (1, 3, 5))
    assert (y - t(g[tag + ".y"])).abs().max() < 5e-6


def test_conv_index_maps():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 40, generator=g, dtype=torch.float64)
    for k, d in ((3, 1), (7, 3), (11, 5)):
        w = torch.randn(5, 6, k, generator=g, dtype=torch.float64)
        b = torch.randn(5, generator=g, dtype=torch.float64)
        assert (O.conv1d(x, w, b, d) - O.conv1d_indexed(x, w, b, d)).abs().max() < 1e-12
    for u in (2, 4):
        w = torch.randn(6, 3, 2 * u, generator=g, dtype=torch.float64)
        b = torch.randn(3, generator=g, dtype=torch.float64)
        y = O.conv_transpose1d(x, w, b, u)
        assert y.shape[-1] == u * 40
        assert (y - O.conv_transpose1d_polyphase(x, w, b, u)).abs().max() < 1e-12


@pytest.mark.parametrize("name", ["tiny", "tiny_tanh_bias", "full"])
def test_generator_vs_reference_golden(golden, synth, cfg, name):
    g = golden("generators")
    h = {"tiny": cfg.tiny_hparams(),
         "tiny_tanh_bias": cfg.tiny_hparams(use_tanh_at_final=True, use_bias_at_final=True),
         "full": cfg.default_hparams()}[name]
    sd = synth.make_state_dict(h, seed=int(g[name + ".seed"][0]))
    # the seeded weights are the ones the reference saw when the golden was made
    assert np.allclose(sd_fingerprint(sd), g[name + ".sd_fingerprint"], rtol=1e-9)
    wav = O.generator_forward(sd, h, t(g[name + ".mel"]))
    ref = t(g[name + ".wav"])
    assert wav.shape == ref.shape
    assert (wav - ref).abs().max() < 1e-5 * max(1.0, float(ref.abs().max()))
    if name == "tiny":
        assert (wav - t(g["tiny.wav_weightnorm"])).abs().max() < 1e-5


def test_fold_weight_norm():
    v = torch.randn(4, 3, 5)
    g = torch.rand(4, 1, 1) + 0.5
    w = O.fold_weight_norm({"c.weight_g": g, "c.weight_v": v, "c.bias": torch.zeros(4)})["c.weight"]
    assert torch.allclose(w.reshape(4, -1).norm(dim=1), g.reshape(-1), atol=1e-6)


def test_survey_counts(synth, cfg):
    h = cfg.default_hparams()
    spec = synth.state_dict_spec(h)
    assert len(spec) == 667
    n = sum(int(np.prod(s)) for k, s, kind in spec if kind != "filter")
    assert n == 112_199_472
    assert cfg.macs_per_frame(h) == 901_859_328
    assert cfg.act_elems_per_frame(h) == 614_400


def test_receptive_field(synth, cfg):
    """Changing one mel frame alters only +-34 frames of output (SURVEY.md 8(c)
    golden fact 4), checked on the tiny generator scaled accordingly."""
    h = cfg.tiny_hparams()
    sd = synth.make_state_dict(h, seed=7)
    mel = synth.make_mel(1, h["num_mels"], 120)
    mel2 = mel.clone()
    mel2[:, :, 60] += 1.0
    up = cfg.total_upsample(h)
    d = (O.generator_forward(sd, h, mel) - O.generator_forward(sd, h, mel2)).abs().reshape(-1)
    nz = torch.nonzero(d > 0).reshape(-1)
    assert int(nz.min()) >= (60 - 34) * up and int(nz.max()) < (60 + 35) * up


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference(synth, cfg):
    h = cfg.tiny_hparams()
    sd = synth.make_state_dict(h, seed=11)
    m = refshim.build_generator(h, sd)
    mel = synth.make_mel(1, h["num_mels"], 13)
    with torch.no_grad():
        ref = m(mel)
    assert (O.generator_forward(sd, h, mel) - ref).abs().max() < 2e-6
    assert set(m.state_dict().keys()) == set(sd.keys())


def test_staged_form_is_the_reference_operator_sequence(synth, cfg):
    """`O.staged_ops()` (the form bench.py times as the CPU arm) runs every Activation1d as the literal F.pad / conv_transpose1d /
    conv1d sequence of the reference: equal to the closed polyphase form to rounding, and - with the reference mounted -
    BIT-identical to the reference's own forward (same torch ops in the same order)."""
    h = cfg.tiny_hparams()
    sd = synth.make_state_dict(h, seed=5)
    mel = synth.make_mel(2, h["num_mels"], 17)
    closed = O.generator_forward(sd, h, mel)
    with O.staged_ops():
        staged = O.generator_forward(sd, h, mel)
    assert not O._STAGED
    assert (staged - closed).abs().max() <= 2e-6
    if refshim.available():
        with torch.no_grad():
            ref = refshim.build_generator(h, sd)(mel)
        assert torch.equal(staged, ref)


# ---- IndexTTS-v1 speaker-conditioned generator ------------------------------------------
V1_CASES = {
    "v1_tiny": lambda cfg: cfg.tiny_v1_hparams(),
    "v1_tiny_nocond_up": lambda cfg: cfg.tiny_v1_hparams(cond_d_vector_in_each_upsampling_layer=False,
                                                         upsample_rates=[4, 2, 2], upsample_kernel_sizes=[4, 2, 4]),
}


@pytest.mark.parametrize("name", sorted(V1_CASES))
def test_v1_generator_vs_reference_golden(golden, synth, cfg, name):
    """the unmodified `indextts.BigVGAN.models.BigVGAN.forward(latent, mel_ref)` (its own ECAPA encoder included) against
    the oracle fed the embedding that encoder produced"""
    g = golden("generators_v1")
    h = V1_CASES[name](cfg)
    sd = synth.make_state_dict(h, seed=int(g[name + ".seed"][0]))
    assert np.allclose(sd_fingerprint(sd), g[name + ".sd_fingerprint"], rtol=1e-12)
    latent, emb, ref = t(g[name + ".latent"]), t(g[name + ".emb"]), t(g[name + ".wav"])
    assert np.array_equal(latent.numpy(), synth.make_latent(latent.shape[0], latent.shape[1], h["gpt_dim"]).numpy())
    wav = O.generator_v1_forward(sd, h, latent, emb)
    assert wav.shape == ref.shape
    assert (wav - ref).abs().max() <= 1e-5 * float(ref.abs().max())
    wav64 = O.generator_v1_forward(sd, h, latent, emb, dtype=torch.float64)
    assert (wav64.float() - ref).abs().max() <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("k,u", [(8, 4), (4, 4), (4, 2), (2, 2), (16, 8), (6, 2), (3, 1), (5, 3)])
def test_conv_transpose_3tap_polyphase_general(k, u):
    """every (k, stride) the native packer accepts is a 3-tap conv over the input rows (layout.cu: pack_convtr_kernel)"""
    g = torch.Generator().manual_seed(k * 100 + u)
    x = torch.randn(2, 5, 9, generator=g, dtype=torch.float64)
    w = torch.randn(5, 3, k, generator=g, dtype=torch.float64)
    b = torch.randn(3, generator=g, dtype=torch.float64)
    ref = O.conv_transpose1d(x, w, b, u)
    assert ref.shape[-1] == 9 * u
    assert (O.conv_transpose1d_taps(x, w, b, u) - ref).abs().max() < 1e-12


# ---- s2mel tail ----------------------------------------------------------------------
S2MEL_CFG_KEYS = ("hidden", "dit_hidden", "n_layers", "kernel_size", "dilation_rate", "out_channels", "freq_dim")


@pytest.mark.parametrize("name", ("full", "ragged", "k3", "k7"))
def test_s2mel_tail_oracle_matches_reference_goldens(golden, synth, name):
    """the oracle restatement against the unmodified reference DiT.forward (goldens of oracle/make_golden.py s2mel)"""
    from oracle import s2mel_oracle as S
    g = golden("s2mel_tail")
    c = dict(zip(S2MEL_CFG_KEYS, (int(v) for v in g[name + ".cfg"])))
    sd = synth.make_s2mel_tail_state_dict(c, seed=int(g[name + ".seed"][0]))
    tt = lambda k: torch.from_numpy(g["%s.%s" % (name, k)])
    ref = tt("out")
    y = S.tail_forward(sd, c, tt("x_res"), tt("x_lens"), tt("t"), tt("t1"))
    assert float((y - ref).abs().max() / ref.abs().max()) <= 1e-5
    y64 = S.tail_forward(sd, c, tt("x_res"), tt("x_lens"), tt("t"), tt("t1"), dtype=torch.float64)
    assert float((y64.float() - ref).abs().max() / ref.abs().max()) <= 1e-5


@pytest.mark.parametrize("name", ("euler_cfg", "euler_nocfg", "euler_cfg2"))
def test_s2mel_euler_oracle_bit_exact_vs_reference_goldens(golden, name):
    from oracle import s2mel_oracle as S
    g = golden("s2mel_tail")
    z, prompt, mu, style, ref = (torch.from_numpy(g["%s.%s" % (name, k)]) for k in ("z", "prompt", "mu", "style", "out"))
    steps, rate = int(g[name + ".meta"][0]), float(g[name + ".meta"][1])
    B, _, T = z.shape
    y = S.solve_euler(S.toy_estimator, z, torch.tensor([T] * B), prompt, mu, style, torch.linspace(0, 1, steps + 1), rate)
    assert torch.equal(y, ref)


def test_s2mel_tail_host_mirror_key_names(synth, cfg):
    """the host mirror exposes the reference's DiT key names for the tail (diffusion_transformer.py:139-157, wavenet.py:119-138,
    encodec.py:124-138,206-208) and its folded / unfolded mappings round-trip"""
    import importlib
    tm = importlib.import_module("voice-tts_b200.s2mel_tail")
    c = cfg.s2mel_tail_config(hidden=32, dit_hidden=32, n_layers=2)
    m = tm.S2MelTail(c, precision="fp32")
    keys = set(m.state_dict().keys())
    for k in ("conv1.weight", "conv1.bias", "t_embedder2.mlp.0.weight", "t_embedder2.mlp.2.bias", "t_embedder2.freqs",
              "wavenet.cond_layer.conv.conv.weight_g", "wavenet.cond_layer.conv.conv.weight_v", "wavenet.cond_layer.conv.conv.bias",
              "wavenet.in_layers.1.conv.conv.weight_v", "wavenet.res_skip_layers.0.conv.conv.bias", "res_projection.weight",
              "final_layer.linear.weight_g", "final_layer.linear.weight_v", "final_layer.adaLN_modulation.1.weight", "conv2.weight"):
        assert k in keys, k
    assert all(k.startswith(tm.TAIL_PREFIXES) for k in keys)
    sd = synth.make_s2mel_tail_state_dict(c, seed=3)
    m.load_folded_state_dict(sd)
    back = m.folded_state_dict()
    assert set(back) == set(sd)
    for k in sd:
        assert torch.allclose(back[k], sd[k], rtol=1e-6, atol=1e-7), k


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_s2mel_tail_takes_the_reference_dit_state_dict(cfg):
    """drop-in check against the REAL reference module: `S2MelTail.from_dit` reads its configuration off a reference DiT and
    loads the tail's share of its state dict strictly (key names and shapes, weight_g / weight_v included); the folded weights
    equal what the reference's weight-normed layers evaluate to, and the oracle on them reproduces DiT.forward"""
    import importlib, sys, types
    if "munch" not in sys.modules:
        sys.modules["munch"] = types.ModuleType("munch")
        sys.modules["munch"].Munch = dict
    refshim.load()
    from indextts.s2mel.modules import diffusion_transformer
    from oracle.make_golden import s2mel_args
    from oracle import s2mel_oracle as S
    tm = importlib.import_module("voice-tts_b200.s2mel_tail")
    c = cfg.s2mel_tail_config(hidden=64, dit_hidden=64, n_layers=3)
    torch.manual_seed(3)
    dit = diffusion_transformer.DiT(s2mel_args(c)).eval()
    tail = tm.S2MelTail.from_dit(dit, precision="fp32")
    assert tail.cfg == dict(c)
    sd = tail.folded_state_dict()
    assert torch.allclose(sd["wavenet.in_layers.1.weight"], dit.wavenet.in_layers[1].conv.conv.weight, atol=1e-7)
    assert torch.allclose(sd["final_layer.linear.weight"], dit.final_layer.linear.weight, atol=1e-7)
    B, T = 2, 23
    dit.setup_caches(B, T)
    cap = {}
    dit.skip_linear.register_forward_hook(lambda m, i, o: cap.__setitem__("x_res", o.detach().clone()))
    dit.t_embedder.register_forward_hook(lambda m, i, o: cap.__setitem__("t1", o.detach().clone()))
    g = torch.Generator().manual_seed(5)
    x_lens = torch.tensor([T, 11])
    tt = torch.rand(B, generator=g)
    with torch.no_grad():
        y = dit(torch.randn(B, 80, T, generator=g), torch.randn(B, 80, T, generator=g), x_lens, tt,
                torch.randn(B, 24, generator=g), torch.randn(B, T, 64, generator=g))
        ours = S.tail_forward({k: v.detach() for k, v in sd.items()}, c, cap["x_res"], x_lens, tt, cap["t1"])
    assert float((ours - y).abs().max() / y.abs().max()) <= 1e-5
