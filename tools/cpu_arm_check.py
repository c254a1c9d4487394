"""CPU arm calibration (build container only: imports the real reference from /root/reference): wall time of one forward of
the unmodified reference `BigVGAN.forward` against the oracle port in its two forms, same weights, same mel, all host threads.
  python tools/cpu_arm_check.py [frames]"""
import importlib, os, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import bigvgan_oracle as O, refshim
cfg = importlib.import_module("voice-tts_b200.config"); synth = importlib.import_module("voice-tts_b200.synth")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 172
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
mel = synth.make_mel(1, 80, frames)
ref = refshim.build_generator(h, sd)
def med(fn, n=3):
    fn(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); y = fn(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[n // 2], y
with torch.no_grad():
    t_ref, y_ref = med(lambda: ref(mel))
    t_closed, y_c = med(lambda: O.generator_forward(sd, h, mel))
    def staged():
        with O.staged_ops():
            return O.generator_forward(sd, h, mel)
    t_staged, y_s = med(staged)
print("threads %d, %d frames: reference %.3f s | oracle closed form %.3f s (%.2fx) | oracle staged %.3f s (%.2fx)" % (
    torch.get_num_threads(), frames, t_ref, t_closed, t_closed / t_ref, t_staged, t_staged / t_ref))
print("max |staged - reference| = %.2e, max |closed - reference| = %.2e (max |ref| %.3f)" % (
    (y_s - y_ref).abs().max(), (y_c - y_ref).abs().max(), y_ref.abs().max()))
