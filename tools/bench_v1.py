"""Throughput of the speaker-conditioned IndexTTS-v1 generator (published plan: 1024-d latent, 1536 initial channels,
x1024 upsampling to 24 kHz, 512-d speaker embedding) - the SURVEY 8(f) rank-2 row measured like the v2 headline:
device-resident inputs, CUDA events, W warm-ups + K steps; prints one JSON line.
usage: python tools/bench_v1.py [B] [T0] [steps]"""
import contextlib, importlib, io, json, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 234          # 234 * 1024 / 24000 = 9.98 s
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
h = cfg.v1_hparams(); sd = synth.make_state_dict(h, 1234)
m = pkg.BigVGANv1(h, precision="bf16")
with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
m.load_state_dict(sd); m = m.to("cuda:0").eval()
x = synth.make_latent(B, T0, h["gpt_dim"]).to("cuda:0"); e = synth.make_speaker_embedding(B, h["speaker_embedding_dim"]).to("cuda:0")
with torch.no_grad():
    for _ in range(3): w, _ = m(x, speaker_embedding=e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): w, _ = m(x, speaker_embedding=e)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    m.set_option("streams", 1)          # serial schedule for the per-kernel split (concurrent streams would double-count time)
    m.set_option("profile", 1); m(x, speaker_embedding=e); m.read_profile(); m(x, speaker_embedding=e); prof = m.read_profile()
audio_s = B * T0 * cfg.total_upsample(h) / h["sampling_rate"]
flops = 2.0 * cfg.macs_per_frame(h) * B * T0
print(json.dumps({"metric": "vocoder_audio_seconds_per_second", "generator": "IndexTTS-v1 speaker-conditioned BigVGAN (published plan)",
                  "value": audio_s / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms, "batch": B, "latent_frames": T0,
                  "audio_s_per_step": audio_s, "launches": m.last_forward_launches(), "dense_tflops_per_step": flops / 1e12,
                  "achieved_tflops_whole_step": flops / (ms * 1e-3) / 1e12, "wav_absmax": float(w.abs().max()),
                  "serial_time_split_ms": {k: round(v[0], 3) for k, v in prof.items()}}))
