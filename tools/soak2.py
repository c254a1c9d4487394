"""Long soak of one shape with separate handles per schedule (the structure of tests/test_gpu_vocoder.py::test_multi_stream_soak...)
  python tools/soak2.py N [B T0]"""
import importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T0 = int(sys.argv[3]) if len(sys.argv) > 3 else 172
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
def make(**opts):
    m = pkg.BigVGAN(h, precision="bf16")
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(sd); m = m.to("cuda:0").eval()
    for k, v in opts.items(): m.set_option(k, v)
    return m
mel = synth.make_mel(B, 80, T0).to("cuda:0")
base = make(streams=1)
with torch.no_grad():
    ref = base(mel).clone()
    again = base(mel)
print("serial repeat identical:", bool(torch.equal(ref, again)), flush=True)
tot = 0
for opts in ({"streams": 1}, {"streams": 3, "conv_own_sm": 1}, {"streams": 3, "conv_own_sm": 0}, {"streams": 3, "conv_own_sm": 0, "graph": 1},
             {"streams": 3, "conv_own_sm": 0, "fuse_unit": 1}):
    m = make(**opts)
    nbad = 0
    with torch.no_grad():
        for i in range(N):
            y = m(mel)
            if not torch.equal(y, ref):
                nbad += 1
                d = (y != ref)
                idx = d.nonzero()
                if nbad <= 5:
                    per_utt = d.view(B, -1).sum(1).tolist()
                    first = idx[0].tolist(); last = idx[-1].tolist()
                    print("  %s iter %d: %d samples differ, per utterance %s, first %s last %s, max|err| %.3e" % (
                        opts, i, int(d.sum()), per_utt, first, last, float((y - ref).abs().max())), flush=True)
    print("%s: %d of %d forwards differ" % (opts, nbad, N), flush=True)
    tot += nbad if "fuse_unit" not in opts else 0
    del m; torch.cuda.empty_cache()
print("SOAK2", "FAILED" if tot else "OK")
