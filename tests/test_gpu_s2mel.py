"""GPU parity of the s2mel tail through the C ABI (`bvg_s2mel_tail_fwd`,
`bvg_cfm_euler_step`): against golden vectors produced by the UNMODIFIED reference (`DiT.forward` with forward hooks on
the tail's inputs, `BASECFM.solve_euler` around a toy estimator - oracle/make_golden.py s2mel) and against the oracle.
Bars: fp32 mode <= 1e-5 of the output's max-abs; bf16 mode (bf16 operands of the dense layers, fp32 accumulation, fp32
residual / skip / LayerNorm) by SNR >= 40 dB; the Euler / CFG update bit-exact (it is fp32 elementwise arithmetic)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O
from oracle import s2mel_oracle as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG_KEYS = ("hidden", "dit_hidden", "n_layers", "kernel_size", "dilation_rate", "out_channels", "freq_dim")
CASES = ("full", "ragged", "k3", "k7")


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def tail_mod():
    return importlib.import_module("voice-tts_b200.s2mel_tail")


def load_case(g, name):
    cfg = dict(zip(CFG_KEYS, (int(v) for v in g[name + ".cfg"])))
    return cfg, int(g[name + ".seed"][0]), t(g[name + ".x_res"]), t(g[name + ".x_lens"]), t(g[name + ".t"]), t(g[name + ".t1"]), \
        t(g[name + ".out"])


def make(tail_mod, synth, cfg, seed, precision):
    m = tail_mod.S2MelTail(cfg, precision=precision)
    m.load_folded_state_dict(synth.make_s2mel_tail_state_dict(cfg, seed=seed))
    return m.to(DEV).eval()


@pytest.mark.parametrize("name", CASES)
def test_tail_fp32_vs_reference_golden(tail_mod, synth, golden, name):
    cfg, seed, x_res, x_lens, tt, t1, ref = load_case(golden("s2mel_tail"), name)
    m = make(tail_mod, synth, cfg, seed, "fp32")
    with torch.no_grad():
        y = m(x_res.to(DEV), x_lens.to(DEV), tt.to(DEV), t1.to(DEV)).cpu()
    assert y.shape == ref.shape
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err


@pytest.mark.parametrize("name", CASES)
def test_tail_bf16_vs_reference_golden(tail_mod, synth, golden, name):
    cfg, seed, x_res, x_lens, tt, t1, ref = load_case(golden("s2mel_tail"), name)
    m = make(tail_mod, synth, cfg, seed, "bf16")
    with torch.no_grad():
        y = m(x_res.to(DEV), x_lens.to(DEV), tt.to(DEV), t1.to(DEV)).cpu()
    snr = O.snr_db(ref, y)
    print("s2mel tail bf16 %s: %.1f dB" % (name, snr))
    assert snr >= 40.0, snr


def test_tail_no_lens_equals_full_lens_and_is_deterministic(tail_mod, synth, cfg):
    c = cfg.s2mel_tail_config(hidden=64, dit_hidden=64, n_layers=3)
    m = make(tail_mod, synth, c, 5, "bf16")
    x_res, tt, t1, lens = synth.make_s2mel_tail_inputs(c, 2, 45)
    with torch.no_grad():
        a = m(x_res.to(DEV), None, tt.to(DEV), t1.to(DEV))
        b = m(x_res.to(DEV), lens.to(DEV), tt.to(DEV), t1.to(DEV))
        c2 = m(x_res.to(DEV), lens.to(DEV), tt.to(DEV), t1.to(DEV))
    assert torch.equal(a, b) and torch.equal(b, c2)


@pytest.mark.parametrize("precision,bar", [("fp32", None), ("bf16", 40.0)])
def test_tail_production_shape_vs_oracle(tail_mod, synth, cfg, precision, bar):
    """the shape infer_v2 runs per solver step (flow_matching.py:88-98: the CFG-stacked batch of 2, here 700 frames), full config"""
    c = cfg.s2mel_tail_config()
    sd = synth.make_s2mel_tail_state_dict(c, seed=99)
    x_res, tt, t1, lens = synth.make_s2mel_tail_inputs(c, 2, 700, seed=3, lens=[700, 512])
    ref = S.tail_forward(sd, c, x_res, lens, tt, t1)
    m = tail_mod.S2MelTail(c, precision=precision)
    m.load_folded_state_dict(sd)
    m = m.to(DEV).eval()
    with torch.no_grad():
        y = m(x_res.to(DEV), lens.to(DEV), tt.to(DEV), t1.to(DEV)).cpu()
    if bar is None:
        assert float((y - ref).abs().max() / ref.abs().max()) <= 1e-5
    else:
        assert O.snr_db(ref, y) >= bar


def test_tail_graph_replay_is_bit_identical(tail_mod, synth, cfg):
    """the launch sequence replays as a CUDA graph from the second forward of a shape on (the solver's 25 estimator calls per
    utterance): same bits as eager launches, also after the shape changes and comes back"""
    c = cfg.s2mel_tail_config(hidden=128, dit_hidden=128, n_layers=3)
    eager = make(tail_mod, synth, c, 9, "bf16")
    eager.set_option("graph", 0)
    graphed = make(tail_mod, synth, c, 9, "bf16")
    with torch.no_grad():
        for T in (50, 50, 50, 81, 50, 81, 81):
            x_res, tt, t1, lens = synth.make_s2mel_tail_inputs(c, 2, T, seed=T)
            lens[1] = T - 7
            args = (x_res.to(DEV), lens.to(DEV), tt.to(DEV), t1.to(DEV))
            assert torch.equal(eager(*args), graphed(*args)), T
    assert graphed.last_forward_launches() == eager.last_forward_launches() > 20


@pytest.mark.parametrize("seed", range(12))
def test_tail_fuzz_vs_oracle(tail_mod, synth, cfg, seed):
    """seeded random configurations (widths off the 64-channel tile grid, 1-4 layers, k = 3 / 5 / 7, ragged lengths incl.
    length-1 and zero-length masks, T down to k // 2 + 1) against the oracle: fp32 mode <= 1e-5, bf16 mode >= 40 dB"""
    import random
    rnd = random.Random(1000 + seed)
    H = rnd.choice([16, 32, 48, 80, 96, 128, 160])
    k = rnd.choice([3, 5, 7])
    c = cfg.s2mel_tail_config(hidden=H, dit_hidden=H, n_layers=rnd.randint(1, 4), kernel_size=k,
                              out_channels=rnd.choice([8, 20, 80, 100]), freq_dim=rnd.choice([32, 256]))
    B = rnd.randint(1, 3)
    T = rnd.choice([k // 2 + 1, k, 9, 17, 40, 63, 90])
    lens = [rnd.choice([0, 1, T // 2, T]) for _ in range(B)]
    lens[rnd.randrange(B)] = T
    sd = synth.make_s2mel_tail_state_dict(c, seed=seed)
    x_res, tt, t1, x_lens = synth.make_s2mel_tail_inputs(c, B, T, seed=seed, lens=lens)
    ref = S.tail_forward(sd, c, x_res, x_lens, tt, t1)
    for precision in ("fp32", "bf16"):
        m = tail_mod.S2MelTail(c, precision=precision)
        m.load_folded_state_dict(sd)
        m = m.to(DEV).eval()
        with torch.no_grad():
            y = m(x_res.to(DEV), x_lens.to(DEV), tt.to(DEV), t1.to(DEV)).cpu()
        if precision == "fp32":
            assert float((y - ref).abs().max() / ref.abs().max()) <= 1e-5, (c, B, T, lens)
        else:
            assert O.snr_db(ref, y) >= 40.0, (c, B, T, lens, O.snr_db(ref, y))


def test_tail_rejects_bad_arguments(tail_mod, synth, cfg):
    c = cfg.s2mel_tail_config(hidden=32, dit_hidden=32, n_layers=2)
    m = make(tail_mod, synth, c, 1, "fp32")
    x_res, tt, t1, _ = synth.make_s2mel_tail_inputs(c, 1, 2)
    with pytest.raises(RuntimeError):        # T = 2 <= (k - 1) / 2: reflect padding undefined (the reference pads with zeros first)
        m(x_res.to(DEV), None, tt.to(DEV), t1.to(DEV))
    x_res, tt, t1, _ = synth.make_s2mel_tail_inputs(c, 1, 8)
    with pytest.raises(RuntimeError):
        m(x_res.to(DEV)[..., :16].contiguous(), None, tt.to(DEV), t1.to(DEV))
    with pytest.raises(RuntimeError):
        m(x_res, None, tt, t1)               # CPU tensors: no fallback


@pytest.mark.parametrize("name", ("euler_cfg", "euler_nocfg", "euler_cfg2"))
def test_solve_euler_bit_exact_vs_reference_golden(tail_mod, golden, name):
    g = golden("s2mel_tail")
    z, prompt, mu, style, ref = (t(g["%s.%s" % (name, k)]) for k in ("z", "prompt", "mu", "style", "out"))
    steps, rate = int(g[name + ".meta"][0]), float(g[name + ".meta"][1])
    B, _, T = z.shape
    y = tail_mod.solve_euler(S.toy_estimator, z.to(DEV), torch.tensor([T] * B, device=DEV), prompt.to(DEV), mu.to(DEV),
                             style.to(DEV), None, torch.linspace(0, 1, steps + 1), inference_cfg_rate=rate)
    assert torch.equal(y.cpu(), ref)


def test_euler_step_matches_torch_ops(tail_mod):
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 80, 333, generator=gen)
    d = torch.randn(4, 80, 333, generator=gen)
    for rate, plen in ((0.7, 40), (0.0, 0), (0.3, 333)):
        dd = d if rate > 0 else d[:2].contiguous()
        a, b = dd[:2], dd[2:] if rate > 0 else None
        dphi = (1.0 + rate) * a - rate * b if rate > 0 else a
        ref = x + torch.tensor(0.04) * dphi
        ref[:, :, :plen] = 0
        y = tail_mod.euler_step_(x.clone().to(DEV), dd.to(DEV), 0.04, rate, plen).cpu()
        assert torch.equal(y, ref)
