"""First-contact GPU probe: each component in isolation, printed diagnostics."""
import importlib, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import bigvgan_oracle as O
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth")
ops = importlib.import_module("voice-tts_b200.ops")
dev = "cuda:0"
what = sys.argv[1] if len(sys.argv) > 1 else "all"
taps = O.kaiser_taps()
tl = taps.tolist()

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

if what in ("all", "act"):
    g = torch.Generator().manual_seed(0)
    for (B, C, T) in ((1, 3, 1), (2, 3, 7), (1, 5, 257), (2, 24, 1000), (1, 4, 8192), (1, 2, 20003), (1, 3, 24000)):
        x = torch.randn(B, C, T, generator=g) * 2; a = torch.randn(C, generator=g) * .5; b = torch.randn(C, generator=g) * .5
        ref = O.activation1d(x.double(), a.double(), b.double(), taps.double(), taps.double()).float()
        for dt in (torch.float32, torch.bfloat16):
            for fast in (False, True):
                y = ops.act1d(x.to(dev).to(dt), a.to(dev), b.to(dev), tl, tl, fast).float().cpu()
                r = ref if dt == torch.float32 else O.activation1d(x.to(dt).double(), a.double(), b.double(), taps.double(), taps.double()).float()
                print("act1d  BCT %-16s %-8s fast=%d  maxabs err %.3e  edges %.3e" % ((B, C, T), str(dt)[6:], fast, (y - r).abs().max(), (y[..., :3] - r[..., :3]).abs().max()))
        xc = x.transpose(1, 2).contiguous()
        for out_bf16 in (False, True):
            y = ops.act1d_cl(xc.to(dev), a.to(dev), b.to(dev), tl, tl, out_bf16, False).float().cpu().transpose(1, 2)
            print("act1d  BTC %-16s out_bf16=%d       maxabs err %.3e" % ((B, C, T), out_bf16, (y - ref).abs().max()))

if what in ("all", "simt"):
    g = torch.Generator().manual_seed(1)
    for (B, Cin, Cout, T, k, d) in ((1, 16, 16, 50, 3, 1), (2, 24, 24, 300, 11, 5), (1, 80, 200, 64, 7, 1), (1, 96, 96, 1000, 7, 3)):
        x = torch.randn(B, Cin, T, generator=g); w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** .5; b = torch.randn(Cout, generator=g)
        ref = O.conv1d(x.double(), w.double(), b.double(), d).float()
        y = ops.conv1d(x.to(dev), w.to(dev), b.to(dev), d, "fp32", 0).cpu()
        print("conv1d simt fp32 %-28s rel err %.3e" % ((B, Cin, Cout, T, k, d), rel(y, ref)))
    for (B, Cin, Cout, T, u) in ((1, 32, 16, 40, 4), (2, 48, 24, 33, 2)):
        x = torch.randn(B, Cin, T, generator=g); w = torch.randn(Cin, Cout, 2 * u, generator=g) / (Cin * 2) ** .5; b = torch.randn(Cout, generator=g)
        ref = O.conv_transpose1d(x.double(), w.double(), b.double(), u).float()
        y = ops.conv_transpose1d(x.to(dev), w.to(dev), b.to(dev), u, "fp32", 0).cpu()
        print("convtr  simt fp32 %-28s rel err %.3e" % ((B, Cin, Cout, T, u), rel(y, ref)))

if what.startswith("umma"):
    variant = int(what[4:] or 0)
    g = torch.Generator().manual_seed(2)
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    for (B, Cin, Cout, T, k, d) in ((1, 64, 128, 256, 1, 1), (1, 64, 128, 256, 3, 1), (1, 64, 128, 300, 3, 1), (2, 128, 256, 700, 7, 3),
                                    (1, 32, 32, 500, 11, 5), (1, 48, 48, 500, 3, 5), (1, 80, 1536, 172, 7, 1), (2, 768, 768, 700, 11, 5)):
        x = bf(torch.randn(B, Cin, T, generator=g)); w = bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** .5); b = torch.randn(Cout, generator=g)
        ref = O.conv1d(x.double(), w.double(), b.double(), d).float()
        t0 = time.time()
        y = ops.conv1d(x.to(dev), w.to(dev), b.to(dev), d, "bf16", variant).cpu()
        print("conv1d umma v%d %-30s rel err %.3e  (%.2fs)" % (variant, (B, Cin, Cout, T, k, d), rel(y, ref), time.time() - t0), flush=True)
    for (B, Cin, Cout, T, u) in ((1, 64, 32, 300, 4), (2, 1536, 768, 172, 4)):
        x = bf(torch.randn(B, Cin, T, generator=g)); w = bf(torch.randn(Cin, Cout, 2 * u, generator=g) / (Cin * 2) ** .5); b = torch.randn(Cout, generator=g)
        ref = O.conv_transpose1d(x.double(), w.double(), b.double(), u).float()
        y = ops.conv_transpose1d(x.to(dev), w.to(dev), b.to(dev), u, "bf16", variant).cpu()
        print("convtr umma v%d %-30s rel err %.3e" % (variant, (B, Cin, Cout, T, u), rel(y, ref)), flush=True)

if what in ("all", "voc32"):
    h = pkg.tiny_hparams(); sd = synth.make_state_dict(h, 7); mel = synth.make_mel(2, h["num_mels"], 21)
    ref = O.generator_forward(sd, h, mel)
    m = pkg.BigVGAN(h, precision="fp32"); m.remove_weight_norm(); m.load_state_dict(sd); m = m.to(dev).eval()
    with torch.no_grad(): y = m(mel.to(dev)).cpu()
    print("vocoder tiny fp32: rel err %.3e  SNR %.1f dB  launches %d" % (rel(y, ref), O.snr_db(ref, y), m.last_forward_launches()))
