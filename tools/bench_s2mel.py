"""Throughput of the s2mel tail on one B200: `bvg_s2mel_tail_fwd` at the shape one solver step of
infer_v2 runs (flow_matching.py:88-98: the CFG-stacked batch of 2, prompt + target frames), both precision modes, the fused
Euler / CFG update, and the reference's operator sequence (oracle, torch CPU, all host threads) on the same shape.
  python tools/bench_s2mel.py [T] [B]     -> one JSON line"""
import importlib, json, os, statistics, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import s2mel_oracle as S
tm = importlib.import_module("voice-tts_b200.s2mel_tail"); synth = importlib.import_module("voice-tts_b200.synth")
cfgm = importlib.import_module("voice-tts_b200.config"); _lib = importlib.import_module("voice-tts_b200._lib")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1863       # 795 prompt + 1068 target frames: the docstring example of flow_matching.py:36
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = cfgm.s2mel_tail_config()
H, D, L, k, C = c["hidden"], c["dit_hidden"], c["n_layers"], c["kernel_size"], c["out_channels"]
macs_row = D * H + L * (H * 2 * H * k) + (L - 1) * (H * 2 * H) + H * H + D * H + H * H + H * C
flops = 2.0 * macs_row * B * T
sd = synth.make_s2mel_tail_state_dict(c, seed=1)
x_res, tt, t1, lens = synth.make_s2mel_tail_inputs(c, B, T)
out = {"workload": "s2mel tail (conv1 + WN x%d + res_projection + FinalLayer + conv2), %d x %d frames" % (L, B, T), "gflop_per_call": flops / 1e9}
dev = "cuda:0"
xr, td, t1d, ld = x_res.to(dev), tt.to(dev), t1.to(dev), lens.to(dev)
ref = None
for prec in ("bf16", "fp32"):
    m = tm.S2MelTail(c, precision=prec); m.load_folded_state_dict(sd); m = m.to(dev).eval()
    with torch.no_grad():
        for _ in range(3): y = m(xr, ld, td, t1d)
        n0 = _lib.launch_count(); y = m(xr, ld, td, t1d); launches = _lib.launch_count() - n0
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): y = m(xr, ld, td, t1d)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 20)
    ms = statistics.median(ts)
    m.set_option("graph", 0)
    with torch.no_grad():
        for _ in range(3): y2 = m(xr, ld, td, t1d)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): y2 = m(xr, ld, td, t1d)
        e1.record(); torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / 20
    assert torch.equal(y, y2)
    out[prec] = {"ms_per_call": ms, "ms_per_call_without_graph_replay": eager_ms, "tflops": flops / (ms * 1e-3) / 1e12, "launches": int(launches), "frames_per_s": B * T / (ms * 1e-3)}
    if prec == "fp32": ref = y.cpu()
    else: yb = y.cpu()
from oracle import bigvgan_oracle as O
out["bf16"]["snr_db_vs_fp32_mode"] = O.snr_db(ref, yb)
# Euler / CFG update of one solver step, [1, 80, T] with the stacked [2, 80, T] estimator output
x = torch.randn(1, C, T, device=dev); d = torch.randn(2, C, T, device=dev)
for _ in range(3): tm.euler_step_(x, d, 0.04, 0.7, 100)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): tm.euler_step_(x, d, 0.04, 0.7, 100)
e1.record(); torch.cuda.synchronize()
out["euler_step_us"] = e0.elapsed_time(e1) / 200 * 1e3
# the reference's operator sequence on the host (oracle: <= 7e-7 from the unmodified reference DiT.forward), same shape
torch.set_num_threads(os.cpu_count())
with torch.no_grad():
    S.tail_forward(sd, c, x_res, lens, tt, t1)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); yo = S.tail_forward(sd, c, x_res, lens, tt, t1); ts.append(time.perf_counter() - t0)
out["cpu_reference_ops"] = {"ms_per_call": statistics.median(ts) * 1e3, "cores": os.cpu_count(), "kind": "port",
                            "max_abs_diff_fp32_mode": float((ref - yo).abs().max())}
out["speedup_bf16_vs_cpu"] = out["cpu_reference_ops"]["ms_per_call"] / out["bf16"]["ms_per_call"]
print(json.dumps(out))
