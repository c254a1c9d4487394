"""Soak test of the multi-stream AMP-block schedule: every 3-stream / 2-stream forward must be bit-identical to the serial one.
usage: python tools/stream_soak.py [repeats]"""
import importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 10
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGAN(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval()
for kv in os.environ.get("BVG_OPTS", "").split(","):   # e.g. BVG_OPTS=fuse_act=0,fuse_res=0
    if "=" in kv: m.set_option(kv.split("=")[0], int(kv.split("=")[1]))
bad = 0
with torch.no_grad():
    for (B, T0) in ((16, 861), (8, 500), (4, 300), (3, 977), (1, 2584), (32, 172), (2, 40)):
        mel = synth.make_mel(B, 80, T0).to("cuda:0")
        m.set_option("streams", 1); m.set_option("graph", 0)
        ref = m(mel).clone(); torch.cuda.synchronize()
        for streams, graph in ((3, 0), (2, 0), (3, 1)):
            m.set_option("streams", streams); m.set_option("graph", graph)
            nd = 0
            for _ in range(R):
                y = m(mel); torch.cuda.synchronize()
                nd += int((y != ref).sum())
            bad += nd
            print("B=%d T0=%d streams=%d graph=%d: %d runs, differing samples %d" % (B, T0, streams, graph, R, nd), flush=True)
print("SOAK", "FAILED" if bad else "OK")
