"""GPU parity of one whole AMPBlock1 unit (bigvgan.py:132-141) as ONE kernel (csrc/amp_unit.cu) against the fp64
oracle composition act -> conv -> act -> conv + residual, and against the layer-by-layer CUDA path."""
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EMPTY = None


@pytest.fixture(scope="module")
def ops():
    import importlib
    return importlib.import_module("voice-tts_b200.ops")


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def make_unit(case, seed_off=0):
    B, C, T, k, d = case
    g = torch.Generator().manual_seed(sum(case) + seed_off)
    x = torch.randn(B, C, T, generator=g)
    w1 = bf(torch.randn(C, C, k, generator=g) / (C * k) ** 0.5)
    w2 = bf(torch.randn(C, C, k, generator=g) / (C * k) ** 0.5)
    b1 = torch.randn(C, generator=g) * 0.3
    b2 = torch.randn(C, generator=g) * 0.3
    al = [torch.randn(C, generator=g) * 0.5 for _ in range(2)]
    be = [torch.randn(C, generator=g) * 0.5 for _ in range(2)]
    return x, w1, b1, w2, b2, al, be


def oracle_unit(x, w1, b1, w2, b2, al, be, d, dtype=torch.float64):
    """xt = a1(x); xt = c1(xt); xt = a2(xt); xt = c2(xt); return the branch xt (bigvgan.py:134-138)."""
    taps = O.kaiser_taps().to(dtype)
    xt = O.activation1d(x.to(dtype), al[0].to(dtype), be[0].to(dtype), taps, taps)
    xt = O.conv1d(xt, w1.to(dtype), b1.to(dtype), d)
    xt = O.activation1d(xt, al[1].to(dtype), be[1].to(dtype), taps, taps)
    return O.conv1d(xt, w2.to(dtype), b2.to(dtype), 1)


def run_unit(ops, x, w1, b1, w2, b2, al, be, d, form, accum=None, scale=1.0, out_bf16=False, precision="bf16"):
    tl = O.kaiser_taps().tolist()
    acc = accum.to(DEV) if accum is not None else torch.empty(0, device=DEV)
    return ops.amp_unit(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), al[0].to(DEV), be[0].to(DEV),
                        al[1].to(DEV), be[1].to(DEV), tl, tl, acc, scale, out_bf16, d, precision, form).cpu()


UNIT_CASES = [  # B, C, T, k, dil  - the three narrow stages of the v2 plan with all (k, dil), lengths on and off the tile grid
    (2, 96, 700, 3, 1), (1, 96, 1000, 7, 3), (1, 96, 640, 11, 5), (2, 96, 223, 11, 1), (1, 96, 225, 7, 5),
    (2, 48, 900, 3, 3), (1, 48, 1100, 7, 1), (2, 48, 500, 11, 5), (1, 48, 224, 3, 1), (1, 48, 449, 11, 3),
    (2, 24, 800, 3, 5), (1, 24, 1500, 7, 3), (3, 24, 400, 11, 5), (1, 24, 176, 3, 1), (1, 24, 177, 11, 1), (1, 24, 353, 7, 5),
    (1, 16, 300, 3, 1), (1, 32, 333, 7, 3), (1, 64, 512, 11, 3), (1, 80, 300, 5, 2), (1, 40, 260, 9, 1), (1, 20, 200, 3, 1),
    (1, 24, 16, 3, 1), (1, 48, 17, 7, 1), (2, 96, 40, 11, 5), (1, 24, 64, 11, 5),
]


@pytest.mark.parametrize("case", UNIT_CASES, ids=[str(c) for c in UNIT_CASES])
def test_amp_unit_one_kernel_vs_oracle(ops, case):
    B, C, T, k, d = case
    x, w1, b1, w2, b2, al, be = make_unit(case)
    branch = oracle_unit(x, w1, b1, w2, b2, al, be, d)
    ref = x.double() + branch
    y = run_unit(ops, x, w1, b1, w2, b2, al, be, d, form=2).double()          # one kernel or fail
    y_lw = run_unit(ops, x, w1, b1, w2, b2, al, be, d, form=1).double()       # layer by layer
    assert y.shape == ref.shape
    scale = float(branch.abs().max())
    err = ((y - x.double()) - branch).abs()
    # two bf16 operand roundings (2^-9 relative each) through two convolutions: well inside 2^-6 of the branch's scale
    assert err.max() <= 2.0 ** -6 * scale, "max err %.3e (scale %.3e) at %s" % (err.max(), scale, (err == err.max()).nonzero()[0].tolist())
    snr = O.snr_db(branch.float(), (y - x.double()).float())
    snr_lw = O.snr_db(branch.float(), (y_lw - x.double()).float())
    assert snr >= 42.0 and snr >= snr_lw - 1.0, (snr, snr_lw)
    # the samples next to the sequence ends (replicate rules of both activations, zero padding of both convs)
    for sl in (slice(0, 12), slice(T - 12, T)):
        e = ((y - x.double()) - branch)[:, :, sl].abs().max()
        assert e <= 2.0 ** -6 * scale, (sl, float(e))


def test_amp_unit_epilogue_forms(ops):
    """(x + branch) * scale + accum, and the bf16-rounded result that feeds the next stage's ConvTranspose1d."""
    case = (2, 48, 700, 7, 3)
    x, w1, b1, w2, b2, al, be = make_unit(case, 3)
    g = torch.Generator().manual_seed(77)
    accum = torch.randn(case[0], case[1], case[2], generator=g)
    branch = oracle_unit(x, w1, b1, w2, b2, al, be, case[4])
    ref = (x.double() + branch) / 3.0 + accum.double()
    y = run_unit(ops, x, w1, b1, w2, b2, al, be, case[4], form=2, accum=accum, scale=1.0 / 3.0).double()
    assert (y - ref).abs().max() <= 2.0 ** -6 * float(branch.abs().max()) / 3.0 + 1e-5
    yb = run_unit(ops, x, w1, b1, w2, b2, al, be, case[4], form=2, accum=accum, scale=1.0 / 3.0, out_bf16=True)
    assert torch.equal(yb, bf(yb))
    assert torch.equal(yb, bf(y.float()))
    # layer-by-layer form of the same epilogue
    y_lw = run_unit(ops, x, w1, b1, w2, b2, al, be, case[4], form=1, accum=accum, scale=1.0 / 3.0).double()
    assert (y_lw - ref).abs().max() <= 2.0 ** -6 * float(branch.abs().max()) / 3.0 + 1e-5


def test_amp_unit_batch_independence(ops):
    case = (3, 24, 900, 11, 3)
    x, w1, b1, w2, b2, al, be = make_unit(case, 5)
    y = run_unit(ops, x, w1, b1, w2, b2, al, be, case[4], form=2)
    for i in range(case[0]):
        yi = run_unit(ops, x[i:i + 1].contiguous(), w1, b1, w2, b2, al, be, case[4], form=2)
        assert torch.equal(yi, y[i:i + 1])


def test_amp_unit_full_size_rows(ops):
    """Stage-4 shape of the benchmark (48 channels, 16 utterances): linearity-free size-independent properties -
    every utterance of a batch of copies gives the same samples, and the first utterance matches a short oracle prefix
    (the receptive field of a unit is 5 + 25 + 5 + 5 rows)."""
    B, C, T, k, d = 16, 48, 110208 // 4, 11, 5
    x, w1, b1, w2, b2, al, be = make_unit((1, C, T, k, d), 11)
    xb = x.expand(B, C, T).contiguous()
    y = run_unit(ops, xb, w1, b1, w2, b2, al, be, d, form=2)
    for i in range(1, B):
        assert torch.equal(y[i], y[0])
    n = 600
    branch = oracle_unit(x[:, :, :n + 64], w1, b1, w2, b2, al, be, d)[:, :, :n]
    got = (y[0:1, :, :n].double() - x[:, :, :n].double())
    assert O.snr_db(branch.float(), got.float()) >= 42.0


def test_amp_unit_fp32_mode_layerwise(ops):
    """fp32 mode has no one-kernel form: the layer-by-layer composition meets the <= 1e-5 bar."""
    case = (1, 24, 300, 7, 3)
    x, w1, b1, w2, b2, al, be = make_unit(case, 9)
    ref = x.double() + oracle_unit(x, w1, b1, w2, b2, al, be, case[4])
    y = run_unit(ops, x, w1, b1, w2, b2, al, be, case[4], form=0, precision="fp32").double()
    assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())
    with pytest.raises(RuntimeError):
        run_unit(ops, x, w1, b1, w2, b2, al, be, case[4], form=2, precision="fp32")


def test_amp_unit_rejects_wide_units(ops):
    case = (1, 192, 300, 3, 1)
    x, w1, b1, w2, b2, al, be = make_unit(case)
    with pytest.raises(RuntimeError):
        run_unit(ops, x, w1, b1, w2, b2, al, be, 1, form=2)
    y = run_unit(ops, x, w1, b1, w2, b2, al, be, 1, form=0)      # falls back to the four layers
    assert y.shape == x.shape


def t_(a):
    import numpy as np
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag", ["c16_k3", "c16_k7", "c8_k11"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ampblock1_dropin_vs_reference_golden(pkg, cfg, golden, tag, precision):
    """The stand-alone `AMPBlock1` drop-in (same constructor / state-dict as bigvgan.py:31-147) on the GPU against outputs of
    the unmodified reference block (tests/golden/ampblock1.npz): fp32 mode <= 1e-5, bf16 mode (one kernel per unit) >= 38 dB
    on these 8-16-channel blocks (little averaging over channels)."""
    g = golden("ampblock1")
    C, k = {"c16_k3": (16, 3), "c16_k7": (16, 7), "c8_k11": (8, 11)}[tag]
    import importlib
    bv = importlib.import_module("voice-tts_b200.bigvgan")
    blk = bv.AMPBlock1(cfg.default_hparams(), C, k, (1, 3, 5), activation="snakebeta")
    blk.remove_weight_norm()
    sd = {key[len(tag) + 4:]: t_(g[key]) for key in g.files if key.startswith(tag + ".sd.")}
    blk.load_state_dict(sd)
    blk = blk.to(DEV).eval()
    x, ref = t_(g[tag + ".x"]), t_(g[tag + ".y"])
    with torch.no_grad():
        y = blk(x.to(DEV), precision=precision).cpu()
    assert y.shape == ref.shape
    if precision == "fp32":
        assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())
    else:
        snr = O.snr_db(ref - x, y - x)      # the residual branches: x passes through in fp32
        print(tag, "bf16 branch SNR %.1f dB" % snr)
        assert snr >= 38.0


def test_activation1d_dropin_in_quantizer_construction(pkg):
    """the other call site of the operator, `Activation1d(activation=SnakeBeta(dim, alpha_logscale=True))`
    as built in indextts/s2mel/modules/quantize.py:95-97 (and the FAcodec decoder blocks): same construction with this
    package's classes, forward on the GPU against the oracle, fp32 / bf16 / fp16 inputs."""
    import importlib
    act_mod = importlib.import_module("voice-tts_b200.activation1d")
    dim = 40
    g = torch.Generator().manual_seed(3)
    m = act_mod.Activation1d(activation=act_mod.SnakeBeta(dim, alpha_logscale=True)).to(DEV)
    with torch.no_grad():
        m.act.alpha.copy_(torch.randn(dim, generator=g) * 0.5)
        m.act.beta.copy_(torch.randn(dim, generator=g) * 0.5)
    x = torch.randn(3, dim, 333, generator=g)
    taps = m.upsample.filter.detach().reshape(-1).cpu().double()
    ref = O.activation1d(x.double(), m.act.alpha.detach().cpu().double(), m.act.beta.detach().cpu().double(), taps, taps)
    with torch.no_grad():
        y = m(x.to(DEV)).cpu().double()
    assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())
    for dt, tol in ((torch.bfloat16, 2.0 ** -7), (torch.float16, 2.0 ** -10)):
        xq = x.to(dt)
        refq = O.activation1d(xq.double(), m.act.alpha.detach().cpu().double(), m.act.beta.detach().cpu().double(), taps, taps)
        with torch.no_grad():
            yq = m(xq.to(DEV))
        assert yq.dtype == dt
        assert (yq.cpu().double() - refq).abs().max() <= tol * float(refq.abs().max())
