"""Serial vs multi-stream AMP-block schedule: bit equality of the waveform and step time.
usage: python tools/stream_probe.py [B] [T0] [steps]"""
import importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 861
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGAN(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval()
mel = synth.make_mel(B, 80, T0).to("cuda:0")
ref = None
with torch.no_grad():
    for streams, graph, prof in ((1, 0, 0), (3, 0, 0), (2, 0, 0), (3, 1, 0), (1, 1, 0), (3, 0, 1), (1, 0, 1)):
        m.set_option("streams", streams); m.set_option("graph", graph); m.set_option("profile", prof)
        for _ in range(3): y = m(mel)
        torch.cuda.synchronize()
        if prof: m.read_profile()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K): y = m(mel)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        if ref is None: ref = y.clone()
        same = bool(torch.equal(ref, y))
        extra = ""
        if prof:
            p = m.read_profile()
            extra = "  sum-of-kernel-events: conv %.2f act %.2f other %.2f ms/step" % (p["conv_tcgen05"][0] / K, p["activation"][0] / K, p["other"][0] / K)
        print("streams=%d graph=%d profile=%d  %.3f ms/step  %.0f x realtime  bit-identical=%s%s" % (
            streams, graph, prof, ms, B * T0 * 256 / 22050 / (ms * 1e-3), same, extra), flush=True)
