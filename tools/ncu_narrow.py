"""Single narrow-stage layers (48 channels, 16 x 110 208 rows = stage 4 of the bench workload) for ncu --set full:
  python tools/ncu_narrow.py conv2 | fused | act
conv2: Conv1d k=7 + residual (conv_umma2_kernel<1,0>); fused: Conv1d k=11 d=5 + Activation1d (conv_umma2a_kernel<0>);
act: stand-alone channels-last activation fp32 -> bf16 (act1d_cl_packed_kernel)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ops = importlib.import_module("voice-tts_b200.ops"); synth = importlib.import_module("voice-tts_b200.synth")
what = sys.argv[1]; C = int(sys.argv[2]) if len(sys.argv) > 2 else 48
B, T = 16, 861 * 6144 // C // 1
T = 861 * (6144 // C)
dev = "cuda:0"; g = torch.Generator().manual_seed(0)
taps = [float(v) for v in synth.kaiser_sinc_filter1d().reshape(-1)]
a = (torch.randn(C, generator=g) * 0.5).to(dev); b = (torch.randn(C, generator=g) * 0.5).to(dev)
bias = torch.randn(C, generator=g).to(dev)
for _ in range(2):
    if what == "conv2":
        x = torch.randn(B, C, T, device=dev); w = (torch.randn(C, C, 7, generator=g) / (C * 7) ** .5).to(dev); res = torch.randn(B, C, T, device=dev)
        y = ops.conv1d_res(x, w, bias, res, torch.empty(0, device=dev), 1.0, False, 1, "bf16", 0)
    elif what == "fused":
        x = torch.randn(B, C, T, device=dev); w = (torch.randn(C, C, 11, generator=g) / (C * 11) ** .5).to(dev)
        y = ops.conv1d_act(x, w, bias, a, b, taps, taps, 5, "bf16", 0)
    else:
        x = torch.randn(B, T, C, device=dev)
        y = ops.act1d_cl(x, a, b, taps, taps, True, True)
    torch.cuda.synchronize()
print("ok", what, tuple(y.shape))
