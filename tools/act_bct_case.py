"""One [B,C,T] activation launch for ncu: python tools/act_bct_case.py C T dtype(fp32|bf16) [fast=1]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ops = importlib.import_module("voice-tts_b200.ops"); synth = importlib.import_module("voice-tts_b200.synth")
C, T = int(sys.argv[1]), int(sys.argv[2]); dt = torch.float32 if sys.argv[3] == "fp32" else torch.bfloat16
fast = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
es = 4 if dt == torch.float32 else 2
B = max(1, -(-(256 << 20) // (C * T * es)))
x = torch.randn(B, C, T, device="cuda").to(dt); a = torch.randn(C, device="cuda") * 0.5; b = torch.randn(C, device="cuda") * 0.5
taps = [float(v) for v in synth.kaiser_sinc_filter1d().reshape(-1)]
for _ in range(3): y = ops.act1d(x, a, b, taps, taps, fast)
torch.cuda.synchronize(); print("ok", B, C, T, float(y.float().abs().max()))
