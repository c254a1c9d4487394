"""Per-layer kernel timings of one generator forward (option profile=1)."""
import importlib, os, sys, warnings, contextlib, io, collections
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 861
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/layers.csv"
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGAN(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval(); m.set_option("profile", 1)
m.set_option("streams", 1)   # per-launch event times only mean something on the serial schedule (override with BVG_OPTS=streams=3)
for kv in os.environ.get("BVG_OPTS", "").split(","):   # e.g. BVG_OPTS=fuse_act=0,fuse_res=2
    if "=" in kv: m.set_option(kv.split("=")[0], int(kv.split("=")[1]))
mel = synth.make_mel(B, 80, T0).to("cuda:0")
with torch.no_grad():
    for _ in range(3): m(mel)
    m.read_profile()
    m(mel)
m.dump_profile(out)
agg = collections.OrderedDict()
for ln in open(out).read().splitlines()[1:]:
    cat, cin, cout, k, dil, rows, ms, work, rate = ln.split(",")
    key = (cat, cin, cout, k, dil) if cat != "2" else (cat, cin, k, dil, "")
    a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += float(ms); a[2] += float(work)
print("cat cin cout k dil | launches  ms   rate(TFLOP/s or GB/s)")
for key, (n, ms, work) in agg.items():
    print("%-28s %3d  %8.3f ms  %8.1f" % (" ".join(key), n, ms, work / (ms * 1e-3) / (1e12 if key[0] in "01" else 1e9)))
