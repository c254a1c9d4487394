"""Victim = stand-alone activation kernel, co-runner = persistent tcgen05 conv on another stream (and vice versa)."""
import importlib, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import bigvgan_oracle as O
ops = importlib.import_module("voice-tts_b200.ops")
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
bf = lambda t: t.to(torch.bfloat16).float()
tl = O.kaiser_taps().tolist()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
# co-runner conv: 384 ch, k = 11 (long persistent kernel)
B, C, T, k, d = 8, 384, 3444 * 4, 11, 1
x = bf(torch.randn(B, C, T, generator=g)).to(dev); w = bf(torch.randn(C, C, k, generator=g) / (C * k) ** .5).to(dev)
b = torch.randn(C, generator=g).to(dev); res = torch.randn(B, C, T, generator=g).to(dev); e = torch.empty(0, device=dev)
conv_ref = ops.conv1d_res(x, w, b, res, e, 1.0, False, d, "bf16", 0)
for (Ba, Ta, Ca) in ((16, 13776, 192), (16, 55104, 48), (16, 220416, 24), (16, 3444, 768)):
    xa = torch.randn(Ba, Ta, Ca, generator=g).to(dev); al = (torch.randn(Ca, generator=g) * .5).to(dev); be = (torch.randn(Ca, generator=g) * .5).to(dev)
    act_ref = ops.act1d_cl(xa, al, be, tl, tl, True, True); torch.cuda.synchronize()
    nda = ndc = 0
    for trial in range(8):
        with torch.cuda.stream(s2):
            yc = ops.conv1d_res(x, w, b, res, e, 1.0, False, d, "bf16", 0)
        with torch.cuda.stream(s1):
            for _ in range(6):
                ya = ops.act1d_cl(xa, al, be, tl, tl, True, True)
                nda += int((ya != act_ref).sum())
        torch.cuda.synchronize()
        ndc += int((yc != conv_ref).sum())
    print("act [%d,%d,%d] beside conv: differing act elements %d, differing conv elements %d" % (Ba, Ta, Ca, nda, ndc), flush=True)
