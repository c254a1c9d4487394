"""One whole-AMP-unit kernel (csrc/amp_unit.cu) at a narrow stage's full size of the bench workload, for ncu / debug builds:
  python tools/ncu_unit.py C k dil [reps]      (16 utterances x 861 * 6144 / C rows)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ops = importlib.import_module("voice-tts_b200.ops"); synth = importlib.import_module("voice-tts_b200.synth")
C = int(sys.argv[1]); k = int(sys.argv[2]); d = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
B, T = 16, 861 * (6144 // C)
dev = "cuda:0"; g = torch.Generator().manual_seed(0)
taps = [float(v) for v in synth.kaiser_sinc_filter1d().reshape(-1)]
al = [(torch.randn(C, generator=g) * 0.5).to(dev) for _ in range(2)]; be = [(torch.randn(C, generator=g) * 0.5).to(dev) for _ in range(2)]
b1 = torch.randn(C, generator=g).to(dev); b2 = torch.randn(C, generator=g).to(dev)
w1 = (torch.randn(C, C, k, generator=g) / (C * k) ** .5).to(dev); w2 = (torch.randn(C, C, k, generator=g) / (C * k) ** .5).to(dev)
x = torch.randn(B, C, T, device=dev)
form = int(os.environ.get("UNIT_FORM", "2"))
for _ in range(reps):
    y = ops.amp_unit(x, w1, b1, w2, b2, al[0], be[0], al[1], be[1], taps, taps, torch.empty(0, device=dev), 1.0, False, d, "bf16", form)
    torch.cuda.synchronize()
print("ok", C, k, d, tuple(y.shape))
