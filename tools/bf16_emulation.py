"""CPU emulation of the bf16 precision plan on the UNMODIFIED reference generator (needs /root/reference): which rounding
costs what?  Conv / ConvTranspose inputs rounded to bf16 (activations), weights rounded to bf16, or both; fp32 accumulation
and fp32 residual stream as in the CUDA path.  One 172-frame utterance (BASELINE configs[0])."""
import importlib, os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import refshim, bigvgan_oracle as O
synth = importlib.import_module("voice-tts_b200.synth"); config = importlib.import_module("voice-tts_b200.config")
torch.set_num_threads(os.cpu_count())
h = config.default_hparams(); sd = synth.make_state_dict(h, seed=1234)
mel = synth.make_mel(1, 80, 172)
def bf(x): return x.to(torch.bfloat16).to(torch.float32)
def run(round_act, round_w):
    m = refshim.build_generator(h, sd)
    convs = [mod for mod in m.modules() if isinstance(mod, (torch.nn.Conv1d, torch.nn.ConvTranspose1d))]
    for c in convs:
        if c is m.conv_post: continue          # conv_post runs in fp32 in the CUDA path
        if round_w:
            with torch.no_grad(): c.weight.copy_(bf(c.weight))
        if round_act:
            c.register_forward_pre_hook(lambda mod, inp: (bf(inp[0]),))
    with torch.no_grad(): return m(mel)
ref = run(False, False)
for name, a, w in (("activations bf16, weights fp32", True, False), ("weights bf16, activations fp32", False, True), ("both bf16 (the CUDA plan)", True, True)):
    y = run(a, w)
    print("%-36s SNR %.2f dB  max-abs rel %.2e" % (name, O.snr_db(ref, y), float((y - ref).abs().max() / ref.abs().max())), flush=True)
