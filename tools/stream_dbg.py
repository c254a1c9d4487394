import importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 300
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGAN(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval()
mel = synth.make_mel(B, 80, T0).to("cuda:0")
outs = {}
with torch.no_grad():
    for tag, streams in (("s1a", 1), ("s3a", 3), ("s3b", 3), ("s3c", 3), ("s1b", 1), ("s2a", 2)):
        m.set_option("streams", streams)
        y = m(mel).clone(); torch.cuda.synchronize()
        outs[tag] = y
        print(tag, "max|y|=%.4f" % y.abs().max().item(), "diff vs s1a = %.3e" % (y - outs["s1a"]).abs().max().item(),
              "nan=%d" % int(torch.isnan(y).sum()),
              "first diff idx:", (y != outs["s1a"]).flatten().nonzero()[:3].flatten().tolist(), "ndiff", int((y != outs["s1a"]).sum()), flush=True)
print("s3a==s3b", torch.equal(outs["s3a"], outs["s3b"]), "s3b==s3c", torch.equal(outs["s3b"], outs["s3c"]))
