"""A/B of handle options on the headline workload (16 x 861 frames, bf16): ms per step, median of 5 x 4 steps each.
  python tools/opt_ab.py "conv_own_sm=0" "graph=1" "conv_own_sm=0,graph=1" ...      (the default set is always measured first)"""
import importlib, os, sys, warnings, contextlib, io, statistics
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
B = int(os.environ.get("AB_B", "16")); T0 = int(os.environ.get("AB_T0", "861"))
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
mel = synth.make_mel(B, 80, T0).to("cuda:0")
sets = [""] + sys.argv[1:]
ref = None
for spec in sets:
    m = pkg.BigVGAN(h, precision="bf16")
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(sd); m = m.to("cuda:0").eval()
    for kv in spec.split(","):
        if "=" in kv: m.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    times = []
    with torch.no_grad():
        for _ in range(3): w = m(mel)
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): w = m(mel)
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 4)
    if ref is None: ref = w.clone()
    print("%-40s median %.3f ms  min %.3f  max %.3f  (%.0f audio-s/s)  identical to default: %s  launches %d" % (
        spec or "default", statistics.median(times), min(times), max(times), B * T0 * 256 / 22050 / (statistics.median(times) * 1e-3),
        bool(torch.equal(w, ref)), m.last_forward_launches()), flush=True)
    del m; torch.cuda.empty_cache()
