"""GPU parity of the dense layers: fp32 SIMT mode and the bf16 tcgen05 mode."""
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import importlib
    return importlib.import_module("voice-tts_b200.ops")


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


CONV_CASES = [  # B, Cin, Cout, T, k, dil
    (1, 16, 16, 50, 3, 1), (2, 24, 24, 300, 11, 5), (1, 80, 200, 64, 7, 1), (1, 12, 12, 77, 7, 3),
    (1, 96, 96, 1000, 7, 3), (2, 48, 48, 513, 3, 5), (1, 192, 192, 256, 11, 1), (1, 1, 5, 9, 3, 1),
    (2, 48, 80, 700, 3, 1), (1, 16, 144, 300, 7, 1), (3, 32, 48, 1031, 11, 3),   # partially filled 128-channel tiles
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv1d(ops, case, precision):
    B, Cin, Cout, T, k, d = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    if precision == "bf16":   # bf16 operands, fp32 accumulate: exact w.r.t. the rounded operands
        x, w = bf(x), bf(w)
    ref = O.conv1d(x.double(), w.double(), b.double(), d)
    y = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), d, precision, 0).cpu().double()
    assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())
    ref2 = O.conv1d_indexed(x.double(), w.double(), b.double(), d)
    assert (ref - ref2).abs().max() < 1e-10


@pytest.mark.parametrize("case", [(1, 32, 16, 40, 4), (2, 48, 24, 33, 2), (1, 256, 128, 300, 4), (1, 24, 12, 50, 2)],
                         ids=str)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_transpose1d(ops, case, precision):
    B, Cin, Cout, T, u = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * u, generator=g) / (Cin * 2) ** 0.5
    b = torch.randn(Cout, generator=g)
    if precision == "bf16":
        x, w = bf(x), bf(w)
    ref = O.conv_transpose1d(x.double(), w.double(), b.double(), u)
    y = ops.conv_transpose1d(x.to(DEV), w.to(DEV), b.to(DEV), u, precision, 0).cpu().double()
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("case", [(2, 48, 24, 33, 4, 4), (1, 192, 96, 130, 4, 4), (1, 32, 16, 40, 2, 2), (1, 24, 12, 21, 6, 2),
                                  (1, 16, 16, 19, 5, 3)], ids=str)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_transpose1d_general_kernel(ops, case, precision):
    """k != 2*stride (padding (k - stride)/2): the k == stride layers of the IndexTTS-v1 plan (upsample_rates [4,4,4,4,2,2],
    kernel sizes [8,8,4,4,4,4], indextts/BigVGAN/models.py:154-161) and other 3-tap polyphase shapes"""
    B, Cin, Cout, T, k, u = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, k, generator=g) / (Cin * 2) ** 0.5
    b = torch.randn(Cout, generator=g)
    if precision == "bf16":
        x, w = bf(x), bf(w)
    ref = O.conv_transpose1d(x.double(), w.double(), b.double(), u)
    y = ops.conv_transpose1d(x.to(DEV), w.to(DEV), b.to(DEV), u, precision, 0).cpu().double()
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())


def test_conv1d_full_size_tcgen05(ops):
    """the largest real layer (768ch, k=11, d=5) at 16 x 10 s would be 1.7 TFLOP for the
    CPU oracle; check one utterance slice against the oracle and the rest through
    batch independence + linearity."""
    g = torch.Generator().manual_seed(0)
    B, C, T, k, d = 4, 768, 3444, 11, 5
    x = bf(torch.randn(B, C, T, generator=g))
    w = bf(torch.randn(C, C, k, generator=g) / (C * k) ** 0.5)
    b = torch.randn(C, generator=g)
    y = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), d, "bf16", 0)
    ref = O.conv1d(x[1:2, :, 1000:1400], w, b, d)[:, :, 30:-30]
    got = y[1:2, :, 1030:1370].cpu()
    assert (got - ref).abs().max() <= 2e-5 * float(ref.abs().max())
    y1 = ops.conv1d(x[2:3].contiguous().to(DEV), w.to(DEV), b.to(DEV), d, "bf16", 0)
    assert torch.equal(y1, y[2:3])
    # fp32 SIMT and tcgen05 agree on identical (bf16-representable) operands
    ys = ops.conv1d(x[:1].contiguous().to(DEV), w.to(DEV), b.to(DEV), d, "fp32", 0)
    assert (ys - y[:1]).abs().max() <= 2e-5 * float(ys.abs().max())


RES_CASES = [  # B, Cin, Cout, T, k, dil, with_accum, out_bf16
    (1, 24, 24, 700, 3, 1, False, False), (2, 24, 24, 1500, 11, 5, True, False), (1, 48, 48, 900, 7, 3, True, True),
    (2, 96, 96, 300, 3, 1, False, True), (1, 192, 192, 530, 7, 1, True, False), (1, 384, 384, 300, 3, 5, True, True),
    (1, 32, 80, 257, 3, 1, False, False), (3, 16, 16, 33, 3, 1, True, False), (1, 48, 48, 1, 3, 1, True, True),
]


@pytest.mark.parametrize("case", RES_CASES, ids=[str(c) for c in RES_CASES])
@pytest.mark.parametrize("variant", [0, 8], ids=["v2", "v1"])
def test_conv1d_residual_epilogue(ops, case, variant):
    """dst = (conv + bias + res) * scale + accum - the AMPBlock1 residual (bigvgan.py:132-141) and the resblock
    mean (:369-375) as the generator issues them; both kernel generations against the fp64 oracle."""
    B, Cin, Cout, T, k, d, with_acc, out_bf16 = case
    g = torch.Generator().manual_seed(sum(int(v) for v in case))
    x = bf(torch.randn(B, Cin, T, generator=g))
    w = bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    res = torch.randn(B, Cout, T, generator=g)
    acc = torch.randn(B, Cout, T, generator=g) if with_acc else torch.empty(0)
    scale = 1.0 / 3.0
    ref = (O.conv1d(x.double(), w.double(), b.double(), d) + res.double()) * float(torch.tensor(scale, dtype=torch.float32))
    if with_acc:
        ref = ref + acc.double()
    y = ops.conv1d_res(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), acc.to(DEV), scale, out_bf16, d, "bf16",
                       variant).cpu().double()
    tol = (2.0 ** -8 if out_bf16 else 1e-5) * float(ref.abs().max())
    assert (y - ref).abs().max() <= tol
    if out_bf16:   # every value is a bf16 number
        assert torch.equal(y.float(), bf(y.float()))


@pytest.mark.parametrize("variant", [0, 8], ids=["v2", "v1"])
def test_conv1d_residual_in_place_and_batch_edges(ops, variant):
    """rows of one utterance never leak into the next one (per-utterance zero padding, clipped stores)."""
    g = torch.Generator().manual_seed(7)
    B, C, T, k, d = 3, 48, 300, 11, 5
    x = bf(torch.randn(B, C, T, generator=g)); w = bf(torch.randn(C, C, k, generator=g) / (C * k) ** 0.5)
    b = torch.randn(C, generator=g); res = torch.randn(B, C, T, generator=g)
    y = ops.conv1d_res(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), torch.empty(0, device=DEV), 1.0, False, d, "bf16", variant)
    for i in range(B):
        yi = ops.conv1d_res(x[i:i + 1].contiguous().to(DEV), w.to(DEV), b.to(DEV), res[i:i + 1].contiguous().to(DEV),
                            torch.empty(0, device=DEV), 1.0, False, d, "bf16", variant)
        assert torch.equal(yi, y[i:i + 1])


def test_conv_errors(ops):
    x = torch.zeros(1, 4, 8, device=DEV)
    with pytest.raises(RuntimeError):
        ops.conv1d(x, torch.zeros(4, 4, 4, device=DEV), torch.zeros(4, device=DEV), 1, "fp32", 0)   # even k
    with pytest.raises(RuntimeError):
        ops.conv1d(x, torch.zeros(4, 4, 3, device=DEV), torch.zeros(4, device=DEV), 1, "fp16", 0)
    with pytest.raises(RuntimeError):
        ops.conv_transpose1d(x, torch.zeros(4, 2, 5, device=DEV), torch.zeros(2, device=DEV), 2, "fp32", 0)  # k - u odd
    with pytest.raises(RuntimeError):
        ops.conv_transpose1d(x, torch.zeros(4, 2, 8, device=DEV), torch.zeros(2, device=DEV), 2, "fp32", 0)  # padding 3 > u: more than 3 taps


ACT_CASES = [  # B, Cin, Cout, T, k, dil   (T chosen around the 240-output tiles / 30-row rounds of the fused kernel)
    (1, 24, 24, 700, 3, 1), (2, 24, 24, 1500, 11, 5), (1, 48, 48, 900, 7, 3), (2, 96, 96, 300, 3, 5),
    (1, 192, 192, 530, 7, 1), (1, 384, 384, 300, 3, 3), (1, 32, 80, 257, 3, 1), (3, 16, 16, 33, 3, 1),
    (1, 48, 48, 1, 3, 1), (2, 64, 64, 2, 7, 1), (1, 128, 128, 5, 3, 1), (1, 40, 40, 6, 3, 1), (1, 24, 24, 7, 11, 1),
    (1, 24, 24, 239, 3, 1), (1, 24, 24, 240, 3, 1), (1, 24, 24, 241, 3, 1), (1, 96, 96, 245, 7, 5), (2, 48, 48, 480, 11, 3),
    (1, 768, 768, 250, 3, 1), (1, 56, 56, 31, 3, 1), (1, 24, 24, 29, 3, 1), (1, 104, 104, 123, 3, 1),
]


@pytest.mark.parametrize("case", ACT_CASES, ids=[str(c) for c in ACT_CASES])
def test_conv1d_fused_activation(ops, case):
    """Activation1d(conv1d(x) + bias) as ONE tcgen05 kernel (bigvgan.py:136-138) against the fp64 oracle composition;
    the fp32 accumulator goes straight through the activation, only the result is rounded to bf16."""
    B, Cin, Cout, T, k, d = case
    g = torch.Generator().manual_seed(sum(case))
    x = bf(torch.randn(B, Cin, T, generator=g))
    w = bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    al = torch.randn(Cout, generator=g) * 0.5
    be = torch.randn(Cout, generator=g) * 0.5
    taps = O.kaiser_taps()
    conv = O.conv1d(x.double(), w.double(), b.double(), d)
    ref = O.activation1d(conv, al.double(), be.double(), taps.double(), taps.double())
    tl = taps.tolist()
    y = ops.conv1d_act(x.to(DEV), w.to(DEV), b.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "bf16", 0).cpu().double()
    assert y.shape == ref.shape
    # bf16 rounding of the result (2^-9 relative) + the fast-cosine snake (~1e-6 of the argument)
    tol = 2.0 ** -8 * float(ref.abs().max())
    err = (y - ref).abs()
    assert err.max() <= tol, "max err %.3e at %s (tol %.3e)" % (err.max(), (err == err.max()).nonzero()[0].tolist(), tol)
    assert torch.equal(y.float(), bf(y.float()))
    # the two-kernel composition (conv -> bf16 -> activation kernel) agrees to within its extra rounding step
    y2 = ops.conv1d_act(x.to(DEV), w.to(DEV), b.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "bf16", 16).cpu().double()
    assert (y2 - ref).abs().max() <= 2.0 ** -6 * float(ref.abs().max())
    assert O.snr_db(ref.float(), y.float()) >= O.snr_db(ref.float(), y2.float()) - 0.5


def test_conv1d_fused_activation_fp32_mode(ops):
    """fp32 mode composes the fp32 conv and the fp32 activation kernel (<= 1e-5 relative)."""
    g = torch.Generator().manual_seed(5)
    B, C, T, k, d = 2, 24, 300, 7, 3
    x = torch.randn(B, C, T, generator=g); w = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b = torch.randn(C, generator=g); al = torch.randn(C, generator=g) * 0.5; be = torch.randn(C, generator=g) * 0.5
    taps = O.kaiser_taps(); tl = taps.tolist()
    ref = O.activation1d(O.conv1d(x.double(), w.double(), b.double(), d), al.double(), be.double(), taps.double(), taps.double())
    y = ops.conv1d_act(x.to(DEV), w.to(DEV), b.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "fp32", 0).cpu().double()
    assert (y - ref).abs().max() <= 1e-5 * float(ref.abs().max())


def test_conv1d_fused_activation_batch_independence(ops):
    g = torch.Generator().manual_seed(9)
    B, C, T, k, d = 3, 48, 500, 11, 5
    x = bf(torch.randn(B, C, T, generator=g)); w = bf(torch.randn(C, C, k, generator=g) / (C * k) ** 0.5)
    b = torch.randn(C, generator=g); al = torch.randn(C, generator=g) * 0.5; be = torch.randn(C, generator=g) * 0.5
    tl = O.kaiser_taps().tolist()
    y = ops.conv1d_act(x.to(DEV), w.to(DEV), b.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "bf16", 0)
    for i in range(B):
        yi = ops.conv1d_act(x[i:i + 1].contiguous().to(DEV), w.to(DEV), b.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "bf16", 0)
        assert torch.equal(yi, y[i:i + 1])


RES_ACT_CASES = [  # B, Cin, Cout, T, k, dil
    (1, 192, 192, 530, 7, 1), (2, 384, 384, 300, 3, 1), (1, 768, 768, 250, 3, 1), (1, 24, 24, 700, 3, 1), (2, 48, 48, 481, 11, 1),
    (1, 96, 96, 1, 3, 1), (1, 64, 64, 5, 7, 1), (1, 128, 128, 239, 3, 1), (1, 32, 80, 257, 3, 1), (3, 16, 16, 33, 3, 1),
]


@pytest.mark.parametrize("case", RES_ACT_CASES, ids=[str(c) for c in RES_ACT_CASES])
def test_conv1d_fused_residual_activation(ops, case):
    """y = conv1d(x) + bias + res and Activation1d(y) from ONE tcgen05 kernel (bigvgan.py:134-139) against the fp64 oracle."""
    B, Cin, Cout, T, k, d = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = bf(torch.randn(B, Cin, T, generator=g))
    w = bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    res = torch.randn(B, Cout, T, generator=g)
    al = torch.randn(Cout, generator=g) * 0.5
    be = torch.randn(Cout, generator=g) * 0.5
    taps = O.kaiser_taps(); tl = taps.tolist()
    yref = O.conv1d(x.double(), w.double(), b.double(), d) + res.double()
    aref = O.activation1d(yref, al.double(), be.double(), taps.double(), taps.double())
    ya, y = ops.conv1d_res_act(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "bf16", 0)
    ya, y = ya.cpu().double(), y.cpu().double()
    assert (y - yref).abs().max() <= 1e-5 * float(yref.abs().max())
    assert (ya - aref).abs().max() <= 2.0 ** -8 * float(aref.abs().max())
    assert torch.equal(ya.float(), bf(ya.float()))
    # y is bit-identical to what the plain residual kernel writes, the activation agrees with the two-kernel composition
    y_plain = ops.conv1d_res(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), torch.empty(0, device=DEV), 1.0, False, d, "bf16", 0)
    assert torch.equal(y_plain.cpu().double(), y)
    ya2, y2 = ops.conv1d_res_act(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "bf16", 16)
    assert torch.equal(y2.cpu().double(), y)
    assert (ya2.cpu().double() - aref).abs().max() <= 2.0 ** -8 * float(aref.abs().max())


def test_conv1d_fused_residual_activation_fp32_mode(ops):
    g = torch.Generator().manual_seed(15)
    B, C, T, k, d = 2, 24, 300, 7, 1
    x = torch.randn(B, C, T, generator=g); w = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b = torch.randn(C, generator=g); res = torch.randn(B, C, T, generator=g)
    al = torch.randn(C, generator=g) * 0.5; be = torch.randn(C, generator=g) * 0.5
    taps = O.kaiser_taps(); tl = taps.tolist()
    yref = O.conv1d(x.double(), w.double(), b.double(), d) + res.double()
    aref = O.activation1d(yref, al.double(), be.double(), taps.double(), taps.double())
    ya, y = ops.conv1d_res_act(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), al.to(DEV), be.to(DEV), tl, tl, d, "fp32", 0)
    assert (y.cpu().double() - yref).abs().max() <= 1e-5 * float(yref.abs().max())
    assert (ya.cpu().double() - aref).abs().max() <= 1e-5 * float(aref.abs().max())


def test_fp32_simt_kernels_bit_identical():
    """the second-generation fp32 SIMT kernel (128 x 128/64 tiles) runs the same FMA chain per output as the first (64 x 64):
    tools/simt_ab.py prints an md5 of every result; BVG_SIMT_V1=1 selects the first kernel (read once per process)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({}, {"BVG_SIMT_V1": "1"}):
        e = dict(os.environ)
        e.pop("BVG_SIMT_V1", None)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "simt_ab.py")], capture_output=True, text=True, env=e,
                           cwd=root, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and outs[0].count("\n") == 4
