"""GPU parity: fused anti-aliased activation (C ABI via the torch custom ops)
against the oracle and the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O
from oracle.make_golden import ACT_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def ops():
    import importlib
    return importlib.import_module("voice-tts_b200.ops")


@pytest.fixture(scope="module")
def act_mod():
    import importlib
    return importlib.import_module("voice-tts_b200.activation1d")


@pytest.mark.parametrize("case", ACT_CASES, ids=[c[0] for c in ACT_CASES])
def test_act1d_golden_fp32(golden, act_mod, case):
    """drop-in Activation1d module vs outputs of the reference's torch Activation1d.
    fp32 kernel mode: <= 1e-5 relative (to the tensor's max) is the north-star bar;
    observed ~3e-7."""
    name, B, C, T, kind, logscale, scale = case
    g = golden("activation1d")
    x, a, b = t(g[name + ".x"]), t(g[name + ".alpha"]), t(g[name + ".beta"])
    cls = act_mod.SnakeBeta if kind == "snakebeta" else act_mod.Snake
    act = cls(C, alpha_logscale=logscale)
    with torch.no_grad():
        act.alpha.copy_(a)
        if kind == "snakebeta":
            act.beta.copy_(b)
    m = act_mod.Activation1d(activation=act).to(DEV)
    y = m(x.to(DEV)).cpu()
    ref = t(g[name + ".y64"]).float()
    assert (y - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max()))
    # fast-sin variant stays within 2e-5 relative as well
    m.fast_sin = True
    y2 = m(x.to(DEV)).cpu()
    assert (y2 - ref).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 2, 3), (3, 2, 5), (2, 3, 6), (1, 7, 129), (2, 4, 4096),
                                   (1, 3, 7809), (1, 2, 7937), (1, 2, 15872), (1, 2, 20003), (2, 2, 31232),
                                   (1, 2, 4352), (1, 2, 4608), (1, 1, 4609), (1, 1, 4864), (1, 2, 7936), (1, 1, 7940), (1, 1, 131075)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_act1d_bct_shapes(ops, shape, dtype):
    """ragged / tiny / tile-boundary lengths, aligned (bulk-copy) and unaligned paths; fp32 tiles longer than ~4 500 samples take
    the in-place kernel (one tile buffer), shorter ones and 16-bit I/O the two-buffer kernel - both sides of that switch, a full
    7 936-sample tile, a tile of 4 samples behind it and a 17-tile row are in the list"""
    B, C, T = shape
    g = torch.Generator().manual_seed(B * 1000 + C * 100 + T)
    x = (torch.randn(B, C, T, generator=g) * 2).to(dtype)
    a, b = torch.randn(C, generator=g) * 0.5, torch.randn(C, generator=g) * 0.5
    taps = O.kaiser_taps()
    ref = O.activation1d(x.double(), a.double(), b.double(), taps.double(), taps.double())
    y = ops.act1d(x.to(DEV), a.to(DEV), b.to(DEV), taps.tolist(), taps.tolist(), False).cpu()
    assert y.dtype == dtype and y.shape == x.shape
    if dtype == torch.float32:
        assert (y.double() - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max()))
    else:  # output rounding to bf16: half an ulp = 2^-9 relative
        assert ((y.double() - ref).abs() <= ref.abs() * 2 ** -8 + 1e-6).all()


def test_act1d_edges_follow_torch_not_reference_kernel(ops):
    """first/last 3 samples: replicate padding of the *activated upsampled* signal
    (torch path), where the reference's own CUDA kernel deviates by up to ~7e-3."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 64, generator=g)
    a, b = torch.randn(4, generator=g) * 0.5, torch.randn(4, generator=g) * 0.5
    taps = O.kaiser_taps()
    ref = O.activation1d_staged(x.double(), a.double(), b.double(), taps.double(), taps.double())
    y = ops.act1d(x.to(DEV), a.to(DEV), b.to(DEV), taps.tolist(), taps.tolist(), False).cpu().double()
    assert (y[..., :3] - ref[..., :3]).abs().max() < 1e-5
    assert (y[..., -3:] - ref[..., -3:]).abs().max() < 1e-5


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 7, 5), (1, 13, 24), (2, 100, 16), (1, 700, 24), (1, 40, 768), (3, 259, 32)])
@pytest.mark.parametrize("io", [("f32", "f32"), ("f32", "bf16"), ("bf16", "bf16")])
def test_act1d_channels_last(ops, shape, io):
    B, T, C = shape
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}
    g = torch.Generator().manual_seed(T * 7 + C)
    x = (torch.randn(B, T, C, generator=g) * 2).to(dt[io[0]])
    a, b = torch.randn(C, generator=g) * 0.5, torch.randn(C, generator=g) * 0.5
    taps = O.kaiser_taps()
    ref = O.activation1d(x.double().transpose(1, 2), a.double(), b.double(), taps.double(), taps.double()).transpose(1, 2)
    y = ops.act1d_cl(x.to(DEV), a.to(DEV), b.to(DEV), taps.tolist(), taps.tolist(), io[1] == "bf16", False).cpu()
    assert y.dtype == dt[io[1]]
    if io[1] == "f32":
        assert (y.double() - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max()))
    else:
        assert ((y.double() - ref).abs() <= ref.abs() * 2 ** -8 + 1e-6).all()


def test_act1d_full_size_properties(ops):
    """BASELINE-size tensor ([8,1536,32768] = 403 M elements would take the oracle minutes):
    size-independent properties instead - unit DC gain incl. edges, linearity when the
    periodic term is switched off (beta -> inf), and batch/channel independence."""
    B, C, T = 4, 1536, 32768
    taps = O.kaiser_taps().tolist()
    a = torch.zeros(C, device=DEV)
    big = torch.full((C,), 60.0, device=DEV)
    x = torch.full((B, C, T), 0.75, device=DEV)
    y = ops.act1d(x, a, big, taps, taps, False)
    assert (y - 0.75).abs().max() < 1e-6
    g = torch.Generator(device=DEV).manual_seed(0)
    x1 = torch.randn(B, C, T, device=DEV, generator=g)
    x2 = torch.randn(B, C, T, device=DEV, generator=g)
    f = lambda z: ops.act1d(z, a, big, taps, taps, False)
    assert (f(x1 + 2 * x2) - f(x1) - 2 * f(x2)).abs().max() < 2e-5
    # one row of the big tensor equals the same row processed alone, with real alpha/beta
    al = torch.randn(C, device=DEV, generator=g) * 0.5
    be = torch.randn(C, device=DEV, generator=g) * 0.5
    yb = ops.act1d(x1, al, be, taps, taps, False)
    row = ops.act1d(x1[2:3, 777:778].contiguous(), al[777:778], be[777:778], taps, taps, False)
    assert torch.equal(yb[2:3, 777:778], row)
    ref = O.activation1d(x1[2:3, 777:778].cpu().double(), al[777:778].cpu().double(), be[777:778].cpu().double(),
                         O.kaiser_taps(dtype=torch.float64), O.kaiser_taps(dtype=torch.float64))
    assert (row.cpu().double() - ref).abs().max() < 1e-5 * float(ref.abs().max())


def test_act1d_errors_and_empty(ops, act_mod):
    taps = O.kaiser_taps().tolist()
    a = torch.zeros(3, device=DEV)
    y = ops.act1d(torch.empty(2, 3, 0, device=DEV), a, a, taps, taps, False)   # seq_len 0: no launch
    assert y.shape == (2, 3, 0)
    with pytest.raises(RuntimeError):
        ops.act1d(torch.zeros(2, 3, 8), a, a, taps, taps, False)                 # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        ops.act1d(torch.zeros(2, 3, 8, device=DEV, dtype=torch.float64), a, a, taps, taps, False)   # fp32 / bf16 / fp16 only
    y16 = ops.act1d(torch.zeros(2, 3, 8, device=DEV, dtype=torch.float16), a, a, taps, taps, False)  # type_shim.h:20-43
    assert y16.dtype == torch.float16 and float(y16.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        ops.act1d(torch.zeros(2, 3, 8, device=DEV), a, a, taps[:6], taps, False)
    with pytest.raises(NotImplementedError):
        act_mod.Activation1d(act_mod.SnakeBeta(3), up_ratio=4)


def test_act1d_fp32_large_arguments(ops):
    """the accurate snake (3-term Cody-Waite reduction mod pi + degree-9 polynomial, csrc/common.cuh sin_mod_pi)
    keeps the <= 1e-5 relative bar when a*u reaches the thousands (large activations times exp(alpha) ~ 30)."""
    g = torch.Generator().manual_seed(11)
    B, C, T = 2, 6, 4096
    x = torch.randn(B, C, T, generator=g) * 30.0
    a = torch.full((C,), 3.4) + torch.randn(C, generator=g) * 0.05      # exp(3.4) = 30
    b = torch.randn(C, generator=g) * 0.5 + 3.0                          # keeps the sin^2 term O(1)
    taps = O.kaiser_taps()
    ref = O.activation1d(x.double(), a.double(), b.double(), taps.double(), taps.double())
    assert float((x.abs().max() * a.exp().max())) > 2000.0
    for layout in ("bct", "cl"):
        if layout == "bct":
            y = ops.act1d(x.to(DEV), a.to(DEV), b.to(DEV), taps.tolist(), taps.tolist(), False).cpu()
        else:
            y = ops.act1d_cl(x.transpose(1, 2).contiguous().to(DEV), a.to(DEV), b.to(DEV), taps.tolist(), taps.tolist(),
                             False, False).cpu().transpose(1, 2)
        assert (y.double() - ref).abs().max() <= 1e-5 * float(ref.abs().max()), layout
