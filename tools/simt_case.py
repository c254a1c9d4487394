"""One fp32 SIMT Conv1d launch for ncu: python tools/simt_case.py [C] [T] [k]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ops = importlib.import_module("voice-tts_b200.ops")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 768; T = int(sys.argv[2]) if len(sys.argv) > 2 else 3444; k = int(sys.argv[3]) if len(sys.argv) > 3 else 11
x = torch.randn(8, C, T, device="cuda"); w = torch.randn(C, C, k, device="cuda") / (C * k) ** .5; b = torch.randn(C, device="cuda")
for _ in range(2): y = ops.conv1d(x, w, b, 1, "fp32", 0)
torch.cuda.synchronize(); print("ok", float(y.abs().max()))
