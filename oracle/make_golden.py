"""ORACLE support — generate tests/golden/*.npz from the REAL reference.

Run in the build container (needs /root/reference):
    python oracle/make_golden.py
Imports the unmodified reference modules (oracle/refshim.py), feeds them the
deterministic synthetic weights / inputs of `voice-tts_b200/synth.py`, and
stores inputs + reference outputs as small fixtures.  The reference ships no
tests or golden vectors, so these files are what pins
the oracle (and through it the CUDA path) to the reference.
"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

synth = importlib.import_module("voice-tts_b200.synth")
config = importlib.import_module("voice-tts_b200.config")
OUT = os.path.join(ROOT, "tests", "golden")

ACT_CASES = [
    # name, B, C, T, kind, logscale, input scale
    ("tiny_t1", 1, 2, 1, "snakebeta", True, 1.0),
    ("tiny_t2", 1, 3, 2, "snakebeta", True, 1.0),
    ("short_t7", 2, 3, 7, "snakebeta", True, 1.0),
    ("odd_t37", 2, 3, 37, "snakebeta", True, 1.0),
    ("c24_t64", 1, 24, 64, "snakebeta", True, 1.0),
    ("c5_t257", 1, 5, 257, "snakebeta", True, 3.0),
    ("snake_t100", 1, 4, 100, "snake", True, 1.0),
    ("linear_t50", 1, 4, 50, "snakebeta", False, 1.0),
    ("c48_t1024", 1, 48, 1024, "snakebeta", True, 2.0),
]


def act_inputs(name, B, C, T, scale):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(B, C, T, generator=g) * scale
    a = torch.randn(C, generator=g) * 0.5
    b = torch.randn(C, generator=g) * 0.5
    return x, a, b


def gen_activation(mod):
    from indextts.s2mel.modules.bigvgan import activations
    from indextts.s2mel.modules.bigvgan.alias_free_activation.torch.act import Activation1d
    from indextts.s2mel.modules.bigvgan.alias_free_activation.torch.filter import kaiser_sinc_filter1d
    out = {"taps": kaiser_sinc_filter1d(0.25, 0.3, 12).reshape(-1).numpy()}
    for name, B, C, T, kind, logscale, scale in ACT_CASES:
        x, a, b = act_inputs(name, B, C, T, scale)
        if kind == "snakebeta":
            act = activations.SnakeBeta(C, alpha_logscale=logscale)
        else:
            act = activations.Snake(C, alpha_logscale=logscale)
        if not logscale:
            a, b = torch.exp(a), torch.exp(b)
        with torch.no_grad():
            act.alpha.copy_(a)
            if kind == "snakebeta":
                act.beta.copy_(b)
        m = Activation1d(activation=act).eval()
        with torch.no_grad():
            y = m(x)
            y64 = m.double()(x.double())
        out[name + ".x"] = x.numpy()
        out[name + ".alpha"] = a.numpy()
        out[name + ".beta"] = (b if kind == "snakebeta" else a).numpy()
        out[name + ".y"] = y.numpy()
        out[name + ".y64"] = y64.numpy()
    np.savez_compressed(os.path.join(OUT, "activation1d.npz"), **out)
    print("activation1d.npz:", len(ACT_CASES), "cases")


def gen_ampblock(mod):
    out = {}
    for C, k, T in ((16, 3, 90), (16, 7, 90), (8, 11, 130)):
        h = mod.AttrDict(dict(config.default_hparams()))
        with contextlib.redirect_stdout(io.StringIO()):
            blk = mod.AMPBlock1(h, C, k, (1, 3, 5), activation="snakebeta")
            blk.remove_weight_norm()
        g = torch.Generator().manual_seed(100 * C + k)
        sd = {}
        for key, v in blk.state_dict().items():
            if key.endswith("filter"):
                sd[key] = v.clone()
            elif "act." in key:
                sd[key] = torch.randn(v.shape, generator=g) * 0.5
            else:
                fan = C * k
                sd[key] = (torch.rand(v.shape, generator=g) * 2 - 1) / fan ** 0.5
        blk.load_state_dict(sd)
        x = torch.randn(2, C, T, generator=g)
        with torch.no_grad():
            y = blk.eval()(x)
        tag = "c%d_k%d" % (C, k)
        out[tag + ".x"] = x.numpy()
        out[tag + ".y"] = y.numpy()
        for key, v in sd.items():
            out[tag + ".sd." + key] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "ampblock1.npz"), **out)
    print("ampblock1.npz")


def sd_fingerprint(sd):
    """Per-state-dict drift detector for the seeded generator."""
    s1 = sum(float(v.double().sum()) for v in sd.values())
    s2 = sum(float(v.double().abs().sum()) for v in sd.values())
    return np.array([s1, s2, float(len(sd))])


def gen_generators(mod):
    out = {}
    cases = [
        ("tiny", config.tiny_hparams(), 7, 2, 21),
        ("tiny_tanh_bias", config.tiny_hparams(use_tanh_at_final=True, use_bias_at_final=True), 8, 1, 9),
        ("full", config.default_hparams(), 1234, 1, 40),
    ]
    for name, h, seed, B, T in cases:
        sd = synth.make_state_dict(h, seed=seed)
        m = refshim.build_generator(h, sd)
        mel = synth.make_mel(B, h["num_mels"], T)
        with torch.no_grad():
            wav = m(mel)
        out[name + ".mel"] = mel.numpy()
        out[name + ".wav"] = wav.numpy()
        out[name + ".sd_fingerprint"] = sd_fingerprint(sd)
        out[name + ".seed"] = np.array([seed])
        print(name, tuple(mel.shape), "->", tuple(wav.shape), "absmax %.4f" % wav.abs().max())
        if name == "tiny":
            # weight-normed checkpoint form of the same model: g = ||v||, v = weight
            with contextlib.redirect_stdout(io.StringIO()):
                m2 = mod.BigVGAN(mod.AttrDict(dict(h)))
            wn = {}
            for key, v in m2.state_dict().items():
                base = key[:-9] if key.endswith((".weight_g", ".weight_v")) else None
                if key.endswith(".weight_v"):
                    wn[key] = sd[base + ".weight"] * 0.5
                elif key.endswith(".weight_g"):
                    w = sd[base + ".weight"]
                    wn[key] = w.reshape(w.shape[0], -1).norm(dim=1).reshape(v.shape)
                else:
                    wn[key] = sd[key]
            m2.load_state_dict(wn)
            with torch.no_grad():
                wav2 = m2.eval()(mel)
            out["tiny.wav_weightnorm"] = wav2.numpy()
            print("  weight-normed form max diff %.2e" % (wav2 - wav).abs().max())
    np.savez_compressed(os.path.join(OUT, "generators.npz"), **out)


def gen_headline(mod):
    """The benchmark's own utterances through the unmodified reference: utterance 3 of BASELINE configs[1] (16 x 861 frames,
    10 s) and utterance 0 of configs[0] (172 frames, 2 s), full 112 M-parameter generator, seed 1234.  The mel is not
    stored: synth.make_mel(1, 80, T, first_utterance=u) regenerates it (a checksum guards against RNG drift)."""
    h = config.default_hparams()
    sd = synth.make_state_dict(h, seed=1234)
    m = refshim.build_generator(h, sd)
    out = {"sd_fingerprint": sd_fingerprint(sd)}
    for name, T, u in (("u861", 861, 3), ("u172", 172, 0)):
        mel = synth.make_mel(1, h["num_mels"], T, first_utterance=u)
        with torch.no_grad():
            wav = m(mel)
        out[name + ".wav"] = wav.numpy()
        out[name + ".utterance"] = np.array([u, T])
        out[name + ".mel_checksum"] = np.array([float(mel.double().sum()), float(mel.double().abs().sum())])
        print(name, tuple(mel.shape), "->", tuple(wav.shape), "absmax %.4f" % wav.abs().max())
    np.savez_compressed(os.path.join(OUT, "headline.npz"), **out)


def gen_v1():
    """IndexTTS-v1 speaker-conditioned generator (indextts/BigVGAN/models.py:130-250), UNMODIFIED forward
    `m(latent, mel_ref)` including its randomly initialised ECAPA-TDNN speaker encoder; the embedding the encoder produced
    is stored next to the waveform because the B200 path takes it as an input (the encoder is out of scope)."""
    from indextts.BigVGAN import models as v1
    out = {}
    cases = [
        ("v1_tiny", config.tiny_v1_hparams(), 21, 2, 13),
        ("v1_tiny_nocond_up", config.tiny_v1_hparams(cond_d_vector_in_each_upsampling_layer=False, upsample_rates=[4, 2, 2],
                                                     upsample_kernel_sizes=[4, 2, 4]), 22, 1, 9),
    ]
    for name, h, seed, B, T in cases:
        sd = synth.make_state_dict(h, seed=seed)
        torch.manual_seed(seed)                      # the speaker encoder keeps its own random init
        with contextlib.redirect_stdout(io.StringIO()):
            m = v1.BigVGAN(v1_attr(h))
            m.remove_weight_norm()
        res = m.load_state_dict(sd, strict=False)
        assert not res.unexpected_keys and all(k.startswith("speaker_encoder.") for k in res.missing_keys), res
        m.eval()
        latent = synth.make_latent(B, T, h["gpt_dim"])
        g = torch.Generator().manual_seed(seed)
        mel_ref = torch.randn(B, 60, h["num_mels"], generator=g)          # [B, frames, num_mels] as infer.py:476 passes it
        with torch.no_grad():
            emb = m.speaker_encoder(mel_ref, None)                        # [B, 1, E]
            wav, loss = m(latent, mel_ref)
        assert loss is None
        out[name + ".latent"] = latent.numpy()
        out[name + ".emb"] = emb.reshape(B, -1).numpy()
        out[name + ".wav"] = wav.numpy()
        out[name + ".sd_fingerprint"] = sd_fingerprint(sd)
        out[name + ".seed"] = np.array([seed])
        print(name, tuple(latent.shape), "->", tuple(wav.shape), "absmax %.4f" % wav.abs().max())
    np.savez_compressed(os.path.join(OUT, "generators_v1.npz"), **out)


def v1_attr(h):
    class H(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    return H(dict(h))


S2MEL_CASES = [
    # name, config overrides, B, T, x_lens (None: all frames valid), seed
    ("full", {}, 2, 61, None, 11),
    ("ragged", {"hidden": 64, "dit_hidden": 64, "n_layers": 3}, 3, 37, [37, 20, 5], 12),
    ("k3", {"hidden": 32, "dit_hidden": 32, "n_layers": 2, "kernel_size": 3}, 1, 9, None, 13),
    ("k7", {"hidden": 48, "dit_hidden": 48, "n_layers": 2, "kernel_size": 7, "out_channels": 20}, 2, 130, [130, 77], 14),
]


def s2mel_args(cfg):
    """the `args` namespace DiT / BASECFM read (diffusion_transformer.py:105-175, flow_matching.py:10-29)"""
    from types import SimpleNamespace as NS
    return NS(DiT=NS(hidden_dim=cfg["dit_hidden"], num_heads=2, depth=1, in_channels=cfg["out_channels"], content_type="discrete",
                     content_codebook_size=16, content_dim=cfg["dit_hidden"], is_causal=False, final_layer_type="wavenet",
                     style_condition=True, class_dropout_prob=0.1, long_skip_connection=True, time_as_token=False,
                     style_as_token=False, uvit_skip_connection=True, block_size=8192, zero_prompt_speech_token=False),
              wavenet=NS(hidden_dim=cfg["hidden"], kernel_size=cfg["kernel_size"], dilation_rate=cfg["dilation_rate"],
                         num_layers=cfg["n_layers"], p_dropout=0.2, style_condition=True),
              style_encoder=NS(dim=24), reg_loss_type="l1", dit_type="DiT")


def gen_s2mel():
    """Tail goldens through the UNMODIFIED reference: DiT.forward runs end to end; forward hooks capture the tail's inputs
    (x_res = output of skip_linear, t1 = output of t_embedder); its return value is the tail's output.  The tail weights
    are the deterministic synthetic ones of synth.make_s2mel_tail_state_dict, loaded through the host mirror's key mapping."""
    import types
    if "munch" not in sys.modules:                       # commons.py imports munch (absent here) for an unrelated helper
        sys.modules["munch"] = types.ModuleType("munch")
        sys.modules["munch"].Munch = dict
    refshim.load()
    from indextts.s2mel.modules import diffusion_transformer, flow_matching
    tail_mod = importlib.import_module("voice-tts_b200.s2mel_tail")
    from oracle import s2mel_oracle
    out = {}
    for name, over, B, T, lens, seed in S2MEL_CASES:
        cfg = config.s2mel_tail_config(**over)
        torch.manual_seed(seed)
        dit = diffusion_transformer.DiT(s2mel_args(cfg)).eval()
        sd = synth.make_s2mel_tail_state_dict(cfg, seed=4321 + seed)
        mirror = tail_mod.S2MelTail(cfg, precision="fp32")
        mirror.load_folded_state_dict(sd)
        res = dit.load_state_dict(mirror.state_dict(), strict=False)     # reference key names incl. weight_g / weight_v
        assert not res.unexpected_keys, res.unexpected_keys
        assert not [k for k in res.missing_keys if k.startswith(tail_mod.TAIL_PREFIXES)], res.missing_keys
        dit.setup_caches(B, T)
        cap = {}
        dit.skip_linear.register_forward_hook(lambda m, i, o: cap.__setitem__("x_res", o.detach().clone()))
        dit.t_embedder.register_forward_hook(lambda m, i, o: cap.__setitem__("t1", o.detach().clone()))
        g = torch.Generator().manual_seed(seed)
        C = cfg["out_channels"]
        x = torch.randn(B, C, T, generator=g)
        prompt_x = torch.randn(B, C, T, generator=g) * 0.5
        x_lens = torch.tensor(lens if lens is not None else [T] * B)
        t = torch.rand(B, generator=g)
        style = torch.randn(B, 24, generator=g)
        cond = torch.randn(B, T, cfg["dit_hidden"], generator=g)
        with torch.no_grad():
            y = dit(x, prompt_x, x_lens, t, style, cond)
        # scale x_res up to O(1) entries? no - keep exactly what the reference's transformer produced
        for k, v in (("x_res", cap["x_res"]), ("t1", cap["t1"]), ("t", t), ("x_lens", x_lens.int()), ("out", y)):
            out["%s.%s" % (name, k)] = v.numpy()
        out[name + ".seed"] = np.array([4321 + seed])
        out[name + ".cfg"] = np.array([cfg[k] for k in ("hidden", "dit_hidden", "n_layers", "kernel_size", "dilation_rate",
                                                        "out_channels", "freq_dim")])
        ours = s2mel_oracle.tail_forward(sd, cfg, cap["x_res"], x_lens if lens is not None else None, t, cap["t1"])
        print(name, tuple(cap["x_res"].shape), "->", tuple(y.shape), "absmax %.4f" % y.abs().max(),
              "oracle max|diff| %.3g" % (ours - y).abs().max())
    # the solver: BASECFM.solve_euler around a toy estimator (exact fp32 arithmetic), with and without CFG
    for name, B, C, T, plen, steps, rate, seed in (("euler_cfg", 2, 80, 50, 17, 25, 0.7, 21), ("euler_nocfg", 1, 20, 33, 0, 6, 0.0, 22),
                                                   ("euler_cfg2", 3, 16, 19, 19, 10, 0.5, 23)):
        cfm = flow_matching.BASECFM(s2mel_args(config.s2mel_tail_config()))
        cfm.estimator = s2mel_oracle.toy_estimator
        g = torch.Generator().manual_seed(seed)
        z = torch.randn(B, C, T, generator=g)
        prompt = torch.randn(B, C, plen, generator=g)
        mu = torch.randn(B, T, 96, generator=g)
        style = torch.randn(B, 8, generator=g)
        t_span = torch.linspace(0, 1, steps + 1)
        import contextlib, io
        with torch.no_grad(), contextlib.redirect_stderr(io.StringIO()):
            y = cfm.solve_euler(z.clone(), torch.tensor([T] * B), prompt, mu.clone(), style, None, t_span, inference_cfg_rate=rate)
        ours = s2mel_oracle.solve_euler(s2mel_oracle.toy_estimator, z.clone(), torch.tensor([T] * B), prompt, mu.clone(), style, t_span, rate)
        for k, v in (("z", z), ("prompt", prompt), ("mu", mu), ("style", style), ("out", y)):
            out["%s.%s" % (name, k)] = v.numpy()
        out[name + ".meta"] = np.array([steps, rate])
        print(name, tuple(z.shape), "steps", steps, "cfg", rate, "absmax %.4f" % y.abs().max(), "oracle bit-identical:",
              bool(torch.equal(ours, y)))
    np.savez_compressed(os.path.join(OUT, "s2mel_tail.npz"), **out)


def main():
    assert refshim.available(), "reference tree not found"
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    mod = refshim.load()
    only = sys.argv[1:]
    if not only or "v2" in only:
        gen_activation(mod)
        gen_ampblock(mod)
        gen_generators(mod)
    if not only or "headline" in only:
        gen_headline(mod)
    if not only or "v1" in only:
        gen_v1()
    if not only or "s2mel" in only:
        gen_s2mel()


if __name__ == "__main__":
    main()
