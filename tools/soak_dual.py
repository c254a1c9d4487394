"""Which kernel's result depends on its co-residents?  Needs the dual-execution build:
  BVG_LIB_NAME=libbvg_dual.so BVG_EXTRA_FLAGS=-DBVG_DUAL python voice-tts_b200/build.py
  BVG_LIB_NAME=libbvg_dual.so python tools/soak_dual.py N [conv_own_sm]"""
import ctypes, importlib, os, sys, warnings, contextlib, io, collections
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
_lib = importlib.import_module("voice-tts_b200._lib")
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
own = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(os.environ.get("SOAK_B", "4")); T0 = int(os.environ.get("SOAK_T0", "172"))
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGAN(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval()
m.set_option("streams", 3); m.set_option("conv_own_sm", own)
for kv in os.environ.get("BVG_OPTS", "").split(","):
    if "=" in kv: m.set_option(kv.split("=")[0], int(kv.split("=")[1]))
mel = synth.make_mel(B, 80, T0).to("cuda:0")
hits = collections.Counter(); nfail = 0
counts = (ctypes.c_uint * 8192)(); tags = (ctypes.c_int * 8192)()
for i in range(N):
    lib.bvg_dual_begin()
    with torch.no_grad(): y = m(mel)
    n = lib.bvg_dual_read(counts, tags, 8192)
    bad = [(j, tags[j], counts[j]) for j in range(n) if counts[j]]
    if bad:
        nfail += 1
        for j, t, c in bad: hits[t] += 1
        if nfail <= 10:
            print("iter %d (records %d): %s" % (i, n, bad), flush=True)
            class E(ctypes.Structure):
                _fields_ = [("idx", ctypes.c_longlong), ("a", ctypes.c_float), ("b", ctypes.c_float), ("r", ctypes.c_float * 7), ("tag", ctypes.c_int), ("pad", ctypes.c_int)]
            buf = (E * 256)()
            k = lib.bvg_dual_log(buf, 256)
            ents = sorted([buf[q] for q in range(k)], key=lambda e: e.idx)
            for e in ents[:24]:
                Cp = {24: 32, 48: 48, 96: 96}.get((e.tag // 1000) % 1000, 0)
                print("   idx %d (row %d ch %d) a %.6f b %.6f a-b %+.6f | res[-3..+3 blocks] %s" % (
                    e.idx, e.idx // Cp if Cp else -1, e.idx % Cp if Cp else -1, e.a, e.b, e.a - e.b, " ".join("%.5f" % v for v in e.r)), flush=True)
print("forwards with a self-inconsistent launch: %d of %d (conv_own_sm=%d)" % (nfail, N, own))
print("by tag (1xxxxxx conv, 2xxxxxx conv+act, 4xxxxxx activation; cin*1000 + k*10 + dil):", dict(hits))
