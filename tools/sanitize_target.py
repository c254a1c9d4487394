"""Small end-to-end workload that touches every kernel family (written for compute-sanitizer memcheck / racecheck /
synccheck; the sanitizer is closed on this GPU pool - exit code 86 - so it currently serves as a quick smoke of all
paths): the tiny v2 generator in both modes
(all kernel families: tcgen05 convs incl. fused epilogues, SIMT convs, channels-last and [B,C,T] activations, conv_post),
the tiny v1 conditioned generator, and the stand-alone operator on ragged shapes.
usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import contextlib, importlib, io, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
ops = importlib.import_module("voice-tts_b200.ops")
dev = "cuda:0"
def build(cls, h, precision, **opts):
    m = cls(h, precision=precision)
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(synth.make_state_dict(h, 7)); m = m.to(dev).eval()
    for k, v in opts.items(): m.set_option(k, v)
    return m
with torch.no_grad():
    h = cfg.tiny_hparams()
    for precision, opts in (("bf16", {}), ("bf16", {"fuse_act": 2, "fuse_res": 2}), ("bf16", {"streams": 1, "graph": 1}), ("fp32", {})):
        m = build(pkg.BigVGAN, h, precision, **opts)
        for B, T in ((2, 21), (1, 3), (3, 40)):
            w = m(synth.make_mel(B, h["num_mels"], T).to(dev))
        torch.cuda.synchronize()
        print("v2", precision, opts, tuple(w.shape), float(w.abs().max()))
    h1 = cfg.tiny_v1_hparams()
    for precision in ("bf16", "fp32"):
        m = build(pkg.BigVGANv1, h1, precision)
        w, _ = m(synth.make_latent(2, 9, h1["gpt_dim"]).to(dev), speaker_embedding=synth.make_speaker_embedding(2, h1["speaker_embedding_dim"]).to(dev))
        torch.cuda.synchronize()
        print("v1", precision, tuple(w.shape), float(w.abs().max()))
    taps = [float(v) for v in synth.kaiser_sinc_filter1d().reshape(-1)]
    for B, C, T, dt in ((1, 3, 1, torch.float32), (2, 5, 37, torch.float32), (1, 24, 1000, torch.bfloat16), (2, 48, 4099, torch.float32)):
        x = torch.randn(B, C, T, device=dev).to(dt)
        y = ops.act1d(x, torch.zeros(C, device=dev), torch.zeros(C, device=dev), taps, taps, True)
        torch.cuda.synchronize()
    print("act ok")
