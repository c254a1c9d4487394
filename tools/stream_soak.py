"""Soak of the multi-stream AMP-block schedule: every multi-stream forward must be bit-identical to the serial one.
usage: python tools/stream_soak.py [repeats]          BVG_OPTS=conv_own_sm=0,... sets handle options for all runs"""
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
base_opts = {}
for kv in os.environ.get("BVG_OPTS", "").split(","):   # e.g. BVG_OPTS=fuse_act=0,fuse_res=0
    if "=" in kv: base_opts[kv.split("=")[0]] = int(kv.split("=")[1])
bad = 0
shapes = ((16, 861), (4, 172), (8, 500), (3, 977), (1, 2584), (32, 172), (2, 40))
if os.environ.get("SOAK_SHAPES"):
    shapes = tuple(tuple(int(v) for v in s.split("x")) for s in os.environ["SOAK_SHAPES"].split(","))
with torch.no_grad():
    for (B, T0) in shapes:
        mel = synth.make_mel(B, 80, T0).to("cuda:0")
        m.set_option("streams", 1); m.set_option("graph", 0); m.set_option("conv_own_sm", 1)
        ref = m(mel).clone(); torch.cuda.synchronize()
        for streams, graph, own in ((3, 0, 1), (3, 0, 0), (2, 0, 0), (3, 1, 0)):
            m.set_option("streams", streams); m.set_option("graph", graph); m.set_option("conv_own_sm", own)
            for k, v in base_opts.items(): m.set_option(k, v)
            nd = 0; nruns = 0; first = None
            for _ in range(R):
                y = m(mel); torch.cuda.synchronize()
                d = (y != ref)
                n = int(d.sum())
                if n:
                    nruns += 1
                    if first is None:
                        idx = d.nonzero()
                        first = "first diff at %s, last at %s, max |err| %.3e" % (idx[0].tolist(), idx[-1].tolist(), float((y - ref).abs().max()))
                nd += n
            bad += nd
            print("B=%d T0=%d streams=%d graph=%d own_sm=%d: %d runs, %d differ, differing samples %d %s" % (B, T0, streams, graph, own, R, nruns, nd, first or ""), flush=True)
print("SOAK", "FAILED" if bad else "OK")
