"""Where does the bf16 mode lose its dB?  SNR of the full generator against the reference goldens (tests/golden/headline.npz:
one 861-frame and one 172-frame utterance through the unmodified reference) under option A/Bs, the per-utterance SNR of the
16 x 861 headline batch against this library's fp32 mode, and the ragged-length / sequence-end numbers behind the test bars."""
import importlib, os, sys, warnings, contextlib, io, json
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import bigvgan_oracle as O
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "headline.npz"))
def make(precision, **opts):
    m = pkg.BigVGAN(h, precision=precision)
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(sd); m = m.to("cuda:0").eval()
    for k, v in opts.items(): m.set_option(k, v)
    return m
res = {}
for name in ("u861", "u172"):
    u, T = [int(v) for v in g[name + ".utterance"]]
    mel = synth.make_mel(1, 80, T, first_utterance=u).to("cuda:0")
    ref = torch.from_numpy(g[name + ".wav"])
    for label, prec, opts in (("fp32", "fp32", {}), ("bf16 default", "bf16", {}), ("bf16 accurate sin (no epilogue fusion)", "bf16", {"fast_sin": 0}),
                              ("bf16 fast sin, no epilogue fusion", "bf16", {"fuse_act": 0, "fuse_res": 0}),
                              ("bf16 all conv1+a2 fused (fp32 m everywhere)", "bf16", {"fuse_act": 2}),
                              ("bf16 all fused", "bf16", {"fuse_act": 2, "fuse_res": 2}),
                              ("bf16 one kernel per narrow AMP unit", "bf16", {"fuse_unit": 1}),
                              ("bf16x3", "bf16x3", {})):
        try:
            m = make(prec, **opts)
            with torch.no_grad(): wav = m(mel).cpu()
            snr = O.snr_db(ref, wav); err = float((wav - ref).abs().max() / ref.abs().max())
            res["%s %s" % (name, label)] = (round(snr, 2), err)
            print("%s %-50s SNR %.2f dB  max-abs rel %.2e" % (name, label, snr, err), flush=True)
            del m; torch.cuda.empty_cache()
        except Exception as e:
            print(name, label, "ERROR", str(e)[:100])
# headline batch: bf16 vs own fp32 mode, per utterance
mel = synth.make_mel(16, 80, 861).to("cuda:0")
m32 = make("fp32"); m16 = make("bf16")
with torch.no_grad(): r = m32(mel).cpu(); w = m16(mel).cpu()
per = [O.snr_db(r[i], w[i]) for i in range(16)]
print("16 x 861 bf16 vs fp32 mode: per-utterance SNR min %.2f median %.2f max %.2f" % (min(per), sorted(per)[8], max(per)))
ref3 = torch.from_numpy(g["u861.wav"])
print("   utterance 3: fp32 mode vs reference golden rel err %.2e; bf16 vs golden %.2f dB" % (float((r[3:4] - ref3).abs().max() / ref3.abs().max()), O.snr_db(ref3, w[3:4])))
res["batch16 per-utterance"] = [round(v, 2) for v in per]
for T0 in (1, 2, 3, 7, 15, 16, 30, 31, 60, 61, 121, 241):
    mel = synth.make_mel(3, 80, T0).to("cuda:0")
    with torch.no_grad(): r = m32(mel).cpu(); w = m16(mel).cpu()
    n = min(512, T0 * 256)
    print("T0=%3d: SNR %.2f  head %.2f  tail %.2f" % (T0, O.snr_db(r, w), O.snr_db(r[..., :n], w[..., :n]), O.snr_db(r[..., -n:], w[..., -n:])), flush=True)
json.dump(res, open("gpurun_out/r2_snr_probe.json", "w"))
