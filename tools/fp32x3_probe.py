import importlib, os, sys, contextlib, io, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
from oracle import bigvgan_oracle as O
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/generators.npz"))
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
mel = torch.from_numpy(g["full.mel"]).cuda(); ref = torch.from_numpy(g["full.wav"])
for impl, terms in ((0, 6), (3, 3), (3, 6), (3, 9)):
    m = pkg.BigVGAN(h, precision="fp32")
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(sd); m = m.cuda().eval(); m.set_option("conv_impl", impl); m.set_option("split_terms", terms)
    with torch.no_grad():
        wav = m(mel).cpu()
        big = synth.make_mel(16, 80, 861).cuda()
        for _ in range(2): m(big)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(big); m(big); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
    print("conv_impl %d terms %d: rel err %.2e  SNR %.1f dB   16 x 10 s: %.1f ms per step = %.0f audio-s/s" % (impl, terms, float((wav - ref).abs().max() / ref.abs().max()), O.snr_db(ref, wav), ms, 16 * 861 * 256 / 22050 / (ms * 1e-3)))
    del m
