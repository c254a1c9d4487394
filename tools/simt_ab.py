import importlib, os, sys, hashlib
sys.path.insert(0, os.getcwd())
import torch
ops = importlib.import_module("voice-tts_b200.ops")
g = torch.Generator().manual_seed(1)
for (B, Ci, Co, T, k, d) in ((2, 768, 768, 700, 11, 5), (1, 96, 96, 3000, 7, 3), (3, 48, 48, 1000, 3, 1), (1, 80, 1536, 300, 7, 1)):
    x = torch.randn(B, Ci, T, generator=g).cuda(); w = (torch.randn(Co, Ci, k, generator=g) / (Ci * k) ** .5).cuda(); b = torch.randn(Co, generator=g).cuda()
    y = ops.conv1d(x, w, b, d, "fp32", 0)
    print((B, Ci, Co, T, k, d), hashlib.md5(y.cpu().numpy().tobytes()).hexdigest())
