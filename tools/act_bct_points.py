"""A few points of the [B,C,T] activation sweep (bvg_act1d_fwd), L2 flushed:  python tools/act_bct_points.py"""
import importlib, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ops = importlib.import_module("voice-tts_b200.ops"); synth = importlib.import_module("voice-tts_b200.synth")
dev = "cuda:0"; taps = synth.kaiser_sinc_filter1d().reshape(-1).tolist()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for dt, fast in ((torch.float32, True), (torch.float32, False), (torch.bfloat16, True)):
    for C, T in ((24, 8192), (192, 8192), (48, 131072), (192, 131072), (768, 131072), (48, 2097152), (192, 2097152)):
        es = 2 if dt == torch.bfloat16 else 4
        B = max(1, -(-(256 << 20) // (C * T * es)))
        x = torch.randn(B, C, T, device=dev).to(dt); a = torch.randn(C, device=dev) * 0.5; b = torch.randn(C, device=dev) * 0.5
        for _ in range(2): ops.act1d(x, a, b, taps, taps, fast)
        ts = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.act1d(x, a, b, taps, taps, fast); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gbps = 2 * B * C * T * es / (ms * 1e-3) / 1e9
        print("%-8s %-8s C=%4d T=%7d B=%3d  %.3f ms  %6.0f GB/s  %.3f of 6547.8" % (str(dt).split(".")[1], "fast" if fast else "accurate", C, T, B, ms, gbps, gbps / 6547.8), flush=True)
        del x
