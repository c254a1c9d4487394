"""CPU tests of the host side: the C-ABI library loads and exports every symbol
include/bvg_b200.h declares (no compute without a GPU), drop-in classes keep the
reference's state-dict layout, sharding logic incl. a world_size-2 gloo run, and
the reference arm of bench.py."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import importlib
    importlib.import_module("__graft_entry__").build()
    _lib = importlib.import_module("voice-tts_b200._lib")
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "bvg_b200.h")).read()
    declared = set(re.findall(r"\b(bvg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS.keys())
    assert lib.bvg_abi_version() == 2


def test_config_struct_matches_header(tmp_path):
    """the ctypes mirror of `bvg_config` (voice-tts_b200/_lib.py) has the size and field offsets a C compiler gives the
    struct in include/bvg_b200.h (a drifted mirror would hand the library a garbled configuration)"""
    import importlib
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib = importlib.import_module("voice-tts_b200._lib")
    names = [n for n, _ in _lib.BvgConfig._fields_]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bvg_b200.h"\nint main(void){printf("%zu", sizeof(bvg_config));'
                   + "".join('printf(" %%zu", offsetof(bvg_config, %s));' % n for n in names) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(_lib.BvgConfig)
    assert vals[1:] == [getattr(_lib.BvgConfig, n).offset for n in names]
    hdr = open(os.path.join(ROOT, "include", "bvg_b200.h")).read()
    body = hdr[hdr.index("typedef struct bvg_config {"):hdr.index("} bvg_config;")]
    assert len(re.findall(r"^\s*int\s+\w+", body, re.M)) == len(names)      # every header field is mirrored


def test_s2mel_config_struct_layout_matches_header(tmp_path):
    """the ctypes mirror of bvg_s2mel_config has the header's size and field offsets (gcc), and the s2mel entry points fail
    loudly without a device / on bad arguments (no CPU fallback)"""
    import importlib
    _lib = importlib.import_module("voice-tts_b200._lib")
    names = [n for n, _ in _lib.S2MelConfig._fields_]
    src = tmp_path / "sz2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bvg_b200.h"\nint main(void){printf("%zu", sizeof(bvg_s2mel_config));'
                   + "".join('printf(" %%zu", offsetof(bvg_s2mel_config, %s));' % n for n in names) + "return 0;}\n")
    exe = tmp_path / "sz2"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(_lib.S2MelConfig)
    assert vals[1:] == [getattr(_lib.S2MelConfig, n).offset for n in names]
    lib = _lib.load()
    c = _lib.S2MelConfig()
    c.hidden, c.dit_hidden, c.n_layers, c.kernel_size, c.dilation_rate, c.out_channels, c.freq_dim, c.mode = 64, 64, 2, 4, 1, 80, 256, 1
    h = ctypes.c_void_p()
    assert lib.bvg_s2mel_tail_create(ctypes.byref(c), ctypes.byref(h)) == -1 and not h.value     # even kernel size: BVG_EINVAL
    c.kernel_size, c.dilation_rate = 5, 2
    assert lib.bvg_s2mel_tail_create(ctypes.byref(c), ctypes.byref(h)) == -1 and b"dilation_rate" in lib.bvg_last_error()
    assert lib.bvg_cfm_euler_step(None, None, 0.1, 0.7, 1, 80, 0, 0, None) == 0                  # T == 0: no-op
    assert lib.bvg_cfm_euler_step(None, None, 0.1, 0.7, 1, 80, 10, 0, None) == -1                # null pointers
    if not torch.cuda.is_available():
        c.dilation_rate = 1
        assert lib.bvg_s2mel_tail_create(ctypes.byref(c), ctypes.byref(h)) in (-5, -4) and not h.value   # no device: no fallback
    tm = importlib.import_module("voice-tts_b200.s2mel_tail")
    cfgm = importlib.import_module("voice-tts_b200.config")
    tail = tm.S2MelTail(cfgm.s2mel_tail_config(hidden=32, dit_hidden=32, n_layers=1), precision="fp32")
    with pytest.raises(RuntimeError):          # CPU tensors: the host mirror has no torch fallback either
        tail(torch.zeros(1, 8, 32), None, torch.zeros(1), torch.zeros(1, 32))
    with pytest.raises(RuntimeError):
        tm.euler_step_(torch.zeros(1, 4, 8), torch.zeros(2, 4, 8), 0.1, 0.7, 0)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_compute_entry_points_fail_loudly_without_gpu():
    import importlib
    _lib = importlib.import_module("voice-tts_b200._lib")
    lib = _lib.load()
    cfg = _lib.BvgConfig()
    cfg.num_mels, cfg.upsample_initial_channel, cfg.num_upsamples, cfg.num_kernels, cfg.num_dilations = 8, 32, 1, 1, 1
    cfg.upsample_rates[0], cfg.upsample_kernel_sizes[0], cfg.resblock_kernel_sizes[0] = 2, 4, 3
    cfg.resblock_dilations[0][0] = 1
    cfg.mode = 1
    h = ctypes.c_void_p()
    rc = lib.bvg_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc in (-5, -4) and not h.value          # BVG_ENODEV / BVG_ECUDA: no fallback
    assert lib.bvg_last_error()
    taps = _lib.taps_array([0.0] * 12)
    buf = (ctypes.c_float * 16)()
    rc = lib.bvg_act1d_fwd(ctypes.addressof(buf), ctypes.addressof(buf) + 32, ctypes.addressof(buf), ctypes.addressof(buf),
                           taps, taps, 1, 1, 4, 0, 0, None)
    assert rc < 0
    # argument validation happens before any device work
    assert lib.bvg_act1d_fwd(None, None, None, None, taps, taps, 1, 1, 4, 7, 0, None) == -2   # BVG_EDTYPE
    assert lib.bvg_act1d_fwd(None, None, None, None, taps, taps, 1, 1, 0, 0, 0, None) == 0    # T == 0: no-op
    ops = importlib.import_module("voice-tts_b200.ops")
    with pytest.raises(RuntimeError):
        ops.act1d(torch.zeros(1, 2, 8), torch.zeros(2), torch.zeros(2), [0.0] * 12, [0.0] * 12, False)


def test_dropin_state_dict_layout(pkg, synth, cfg):
    h = cfg.default_hparams()
    names = [k for k, _, _ in synth.state_dict_spec(h)]
    tiny = cfg.tiny_hparams()
    m = pkg.BigVGAN(tiny)
    wn_keys = set(m.state_dict().keys())
    assert "conv_pre.weight_g" in wn_keys and "resblocks.0.convs1.0.weight_v" in wn_keys
    m.remove_weight_norm()
    m.remove_weight_norm()   # idempotent like the reference (bigvgan.py:388-400)
    sd = synth.make_state_dict(tiny, 1)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    a = m.resblocks[0].activations[0]
    assert hasattr(a, "act") and hasattr(a.upsample, "filter") and hasattr(a.downsample.lowpass, "filter")
    assert len(names) == 667
    f = m.folded_state_dict()
    assert all(torch.equal(f[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        pkg.BigVGAN(cfg.tiny_hparams(resblock="2"))


def test_v1_dropin_state_dict_layout(pkg, synth, cfg):
    """`BigVGANv1` carries the reference's v1 key names (indextts/BigVGAN/models.py:149-209): the v2 keys plus
    cond_layer.* / conds.{i}.*; an injected speaker encoder lives under speaker_encoder.* and is not sent to the native plan"""
    h = cfg.tiny_v1_hparams()
    enc = torch.nn.Linear(3, 2)
    m = pkg.BigVGANv1(h, speaker_encoder=enc)
    m.remove_weight_norm()
    sd = synth.make_state_dict(h, 1)
    keys = set(m.state_dict().keys())
    assert {"speaker_encoder.weight", "speaker_encoder.bias"} <= keys
    assert keys - {"speaker_encoder.weight", "speaker_encoder.bias"} == set(sd.keys())
    assert {"cond_layer.weight", "conds.2.bias", "conv_post.bias"} <= set(sd.keys())
    assert tuple(sd["conv_pre.weight"].shape) == (h["upsample_initial_channel"], h["gpt_dim"], 7)
    assert tuple(sd["ups.1.0.weight"].shape) == (48, 24, 4)           # a k == stride stage
    assert not any(k.startswith("speaker_encoder.") for k in m.folded_state_dict())
    assert m._native_conditioning() == (1, h["speaker_embedding_dim"], 1)
    with pytest.raises(ValueError):
        pkg.BigVGANv1(cfg.tiny_hparams())                            # v2 hyper-parameters: no gpt_dim
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, h["gpt_dim"]), speaker_embedding=torch.zeros(1, h["speaker_embedding_dim"]))   # CPU tensor


def test_shard_ranges_and_chunks():
    import importlib
    shard = importlib.import_module("voice-tts_b200.shard")
    for n in (0, 1, 7, 16, 256):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                covered += list(range(lo, hi))
                assert hi - lo in (n // world, n // world + 1)
            assert covered == list(range(n))
    chunks = shard.split_chunks(2584, 600)
    assert chunks[0][:2] == (0, 634) and chunks[-1][1] == 2584
    assert sum(k1 - k0 for _, _, k0, k1 in chunks) == 2584


def test_gloo_world2_sharding_and_timing_reduce(tmp_path):
    """world_size-2 gloo run of the sharding/timing logic bench.py uses under torchrun."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, importlib, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "shard = importlib.import_module('voice-tts_b200.shard')\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = shard.shard_range(16 * w, r, w)\n"
        "mel = torch.arange(16 * w).float().view(-1, 1, 1).repeat(1, 2, 3)\n"
        "lo2, hi2, out = shard.vocode_sharded(lambda m: m.sum(dim=(1, 2), keepdim=True), mel, r, w)\n"
        "assert (lo, hi) == (lo2, hi2) == (16 * r, 16 * r + 16)\n"
        "got = [torch.zeros_like(out) for _ in range(w)]\n"
        "dist.all_gather(got, out)\n"
        "assert torch.equal(torch.cat(got).view(-1), torch.arange(16 * w).float() * 6)\n"
        "m = shard.max_over_ranks(10.0 + r)\n"
        "assert m == 10.0 + (w - 1)\n"
        "dist.barrier(); sys.stdout.write('ok' + str(r) + chr(10)); sys.stdout.flush()\n" % ROOT)
    import socket
    with socket.socket() as sk:          # a free port: fixed ports collide when suites run back to back
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert "ok0" in out.stdout and "ok1" in out.stdout   # one write() per rank: no interleaving


def test_bench_reference_arm_prints_contract_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--frames", "22"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_arms_share_config_and_metric():
    """the reference arm reports on the SAME `config`, metric and unit as the GPU arm (it times a bounded sample of it), for both
    workloads; `--workload c4` is the 256 x 30 s strong-scaling case of BASELINE configs[3]"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import argparse
    for wl, batch, frames in (("headline", 16, 861), ("c4", 256, 2584)):
        a = argparse.Namespace(workload=wl, batch=batch, frames=frames)
        c1, c8 = bench.workload_config(a, 1), bench.workload_config(a, 8)
        assert c1["workload"] == c8["workload"] and "model" not in c1
        assert c1["global_batch"] == batch and c8["global_batch"] == (batch if wl == "c4" else 8 * batch)
    rows = [{"impl": "ours", "dtype": "bfloat16", "snake": "fast", "T": T, "frac_hbm": f, "GBps": 6547.8 * f}
            for T, f in ((8192, 0.3), (131072, 0.4), (2097152, 0.5))]
    s = bench.act_sweep_summary(rows, {"hbm_gbs": 6547.8})["series"]["ours bfloat16 fast"]
    assert s["frac_min"] == 0.3 and s["frac_max"] == 0.5 and s["frac_median"] == 0.4 and s["frac_median_T_ge_131072"] == 0.5


@pytest.mark.parametrize("C,k,d,F", [(24, 11, 1, 4), (24, 7, 1, 4), (48, 11, 1, 2), (12, 7, 1, 8), (24, 11, 3, 4), (16, 5, 2, 8)])
def test_time_folding_identity(C, k, d, F):
    """The identity behind `vocoder.cu: build_fold_twin`: a zero-padded Conv1d over [T, C] equals a
    dilation-1 conv with kf = 2*floor((c*d + F - 1)/F) + 1 taps over the same memory viewed as [T/F, F*C], with
    Wf[tau][(po,co)][(pi,ci)] = W[co][ci][c + (F*tau + pi - po)/d] (0 where undefined).  Checked in float64 on the CPU."""
    import numpy as np
    rng = np.random.default_rng(C * 100 + k * 10 + d)
    T = 8 * F * 3
    W = rng.standard_normal((C, C, k))
    x = rng.standard_normal((T, C))
    cen = (k - 1) // 2
    S = cen * d
    y = np.zeros((T, C))
    for t in range(T):
        for j in range(k):
            tt = t + (j - cen) * d
            if 0 <= tt < T:
                y[t] += W[:, :, j] @ x[tt]
    kf = 2 * ((S + F - 1) // F) + 1
    cf = (kf - 1) // 2
    Wf = np.zeros((kf, F * C, F * C))
    for po in range(F):
        for pi in range(F):
            for tau in range(-cf, cf + 1):
                delta = F * tau + pi - po
                if delta % d:
                    continue
                j = cen + delta // d
                if 0 <= j < k:
                    Wf[tau + cf, po * C:(po + 1) * C, pi * C:(pi + 1) * C] = W[:, :, j]
    xf = x.reshape(T // F, F * C)
    yf = np.zeros((T // F, F * C))
    for t in range(T // F):
        for tau in range(-cf, cf + 1):
            if 0 <= t + tau < T // F:
                yf[t] += Wf[tau + cf] @ xf[t + tau]
    assert np.allclose(yf.reshape(T, C), y, rtol=1e-12, atol=1e-12)
