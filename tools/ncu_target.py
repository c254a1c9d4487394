"""Exactly one generator forward (no warm-up) - the target of ncu captures.
usage: python tools/ncu_target.py [B] [T0]"""
import importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 861
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGAN(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval()
mel = synth.make_mel(B, 80, T0).to("cuda:0")
with torch.no_grad():
    w = m(mel)
torch.cuda.synchronize()
print("ok", tuple(w.shape), float(w.abs().max()))
