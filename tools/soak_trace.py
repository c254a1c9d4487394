"""Find the first launch whose output differs between the serial schedule and the 3-stream / co-resident one.
Needs the trace build:  BVG_LIB_NAME=libbvg_trace.so BVG_EXTRA_FLAGS=-DBVG_TRACE python voice-tts_b200/build.py
  BVG_LIB_NAME=libbvg_trace.so python tools/soak_trace.py N"""
import ctypes, importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
_lib = importlib.import_module("voice-tts_b200._lib")
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
B = int(os.environ.get("SOAK_B", "4")); T0 = int(os.environ.get("SOAK_T0", "172"))
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
def make(**opts):
    m = pkg.BigVGAN(h, precision="bf16")
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(sd); m = m.to("cuda:0").eval()
    for k, v in opts.items(): m.set_option(k, v)
    return m
def traced(m, mel):
    lib.bvg_trace_begin()
    with torch.no_grad(): y = m(mel)
    sums = (ctypes.c_ulonglong * 8192)(); tags = (ctypes.c_int * 8192)()
    n = lib.bvg_trace_read(sums, tags, 8192)
    return y, list(sums[:n]), list(tags[:n])
mel = synth.make_mel(B, 80, T0).to("cuda:0")
base = make(streams=1)
ref, rs, rt = traced(base, mel)
_, rs2, _ = traced(base, mel)
print("launch records per forward:", len(rs), " serial repeat identical:", rs == rs2, flush=True)
m = make(streams=3, conv_own_sm=0)
found = 0
for i in range(N):
    y, s, t = traced(m, mel)
    bad = [j for j in range(len(rs)) if s[j] != rs[j]]
    if bad or not torch.equal(y, ref):
        found += 1
        j = bad[0] if bad else -1
        print("iter %d: %d records differ; first at launch record %d tag %d (previous tags %s); output equal: %s" % (
            i, len(bad), j, t[j] if j >= 0 else -1, t[max(0, j - 3):j], bool(torch.equal(y, ref))), flush=True)
        if found >= 8: break
print("done, %d failing forwards in %d" % (found, i + 1))
