"""GPU parity of the speaker-conditioned IndexTTS-v1 generator (`BigVGANv1`, the drop-in for
indextts/BigVGAN/models.py::BigVGAN as infer.py:476,646 call it) through `bvg_vocoder_fwd_cond`, against golden vectors
produced by the unmodified reference forward and against the oracle.  Same bars as the v2 generator: fp32 kernel mode
<= 1e-5 of the waveform's max-abs; bf16 mode by SNR (tiny random-weight generators: 33 dB)."""
import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.asarray(a))


def make(pkg, h, sd, precision):
    m = pkg.BigVGANv1(h, precision=precision)
    m.remove_weight_norm()
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to(DEV).eval()


CASES = {
    "v1_tiny": lambda cfg: cfg.tiny_v1_hparams(),
    "v1_tiny_nocond_up": lambda cfg: cfg.tiny_v1_hparams(cond_d_vector_in_each_upsampling_layer=False,
                                                         upsample_rates=[4, 2, 2], upsample_kernel_sizes=[4, 2, 4]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_v1_generator_vs_reference_golden(pkg, synth, cfg, golden, name):
    g = golden("generators_v1")
    h = CASES[name](cfg)
    sd = synth.make_state_dict(h, seed=int(g[name + ".seed"][0]))
    latent, emb, ref = t(g[name + ".latent"]), t(g[name + ".emb"]), t(g[name + ".wav"])
    m = make(pkg, h, sd, "fp32")
    with torch.no_grad():
        wav, loss = m(latent.to(DEV), speaker_embedding=emb.to(DEV))
    assert loss is None and wav.shape == ref.shape
    assert (wav.cpu() - ref).abs().max() <= 1e-5 * float(ref.abs().max())
    mb = make(pkg, h, sd, "bf16")
    with torch.no_grad():
        wavb, _ = mb(latent.to(DEV), speaker_embedding=emb.to(DEV))
    snr = O.snr_db(ref, wavb.cpu())
    print("%s bf16 SNR %.1f dB" % (name, snr))
    assert snr >= 33.0
    mx = make(pkg, h, sd, "bf16x3")          # per-utterance bias rows ride on the last term-pair pass
    with torch.no_grad():
        wavx, _ = mx(latent.to(DEV), speaker_embedding=emb.to(DEV))
    snrx = O.snr_db(ref, wavx.cpu())
    print("%s bf16x3 SNR %.1f dB" % (name, snrx))
    assert snrx >= 70.0


def test_v1_speaker_encoder_module_and_batch_rows(pkg, synth, cfg):
    """`forward(x, mel_ref)` with an injected speaker-encoder module (the reference call form, infer.py:476); every
    utterance gets ITS OWN conditioning vector (per-utterance bias rows), checked against per-row oracle runs"""
    h = cfg.tiny_v1_hparams()
    sd = synth.make_state_dict(h, seed=5)
    B, T, E = 3, 17, h["speaker_embedding_dim"]
    latent = synth.make_latent(B, T, h["gpt_dim"])
    emb = synth.make_speaker_embedding(B, E)

    class Enc(torch.nn.Module):                    # stands in for ECAPA_TDNN: [B, frames, mels] -> [B, 1, E]
        def forward(self, mel_ref, lens=None):
            return emb.to(mel_ref.device).unsqueeze(1)

    m = pkg.BigVGANv1(h, precision="fp32", speaker_encoder=Enc())
    m.remove_weight_norm()
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    with torch.no_grad():
        wav, _ = m(latent.to(DEV), torch.zeros(B, 60, h["num_mels"], device=DEV))
    wav = wav.cpu()
    for b in range(B):
        ref = O.generator_v1_forward(sd, h, latent[b:b + 1], emb[b:b + 1])
        assert (wav[b:b + 1] - ref).abs().max() <= 1e-5 * float(ref.abs().max()), b
    # swapping two speakers' embeddings changes exactly those rows
    perm = torch.tensor([1, 0, 2])
    with torch.no_grad():
        wav2, _ = m(latent.to(DEV), speaker_embedding=emb[perm].to(DEV))
    wav2 = wav2.cpu()
    assert torch.equal(wav2[2], wav[2]) and not torch.equal(wav2[0], wav[0])


def test_v1_published_plan_shapes_bf16(pkg, synth, cfg):
    """the published IndexTTS-1 plan (1024-d latent, x1024 upsampling, k = u stages, 512-d speaker embedding) at reduced
    width (initial channels 384) against the oracle: bf16 SNR and fp32 tolerance"""
    h = cfg.v1_hparams(upsample_initial_channel=384)
    sd = synth.make_state_dict(h, seed=3)
    latent = synth.make_latent(2, 6, h["gpt_dim"])
    emb = synth.make_speaker_embedding(2, h["speaker_embedding_dim"])
    ref = O.generator_v1_forward(sd, h, latent, emb)
    assert ref.shape == (2, 1, 6 * 1024)
    for precision in ("fp32", "bf16"):
        m = make(pkg, h, sd, precision)
        with torch.no_grad():
            wav, _ = m(latent.to(DEV), speaker_embedding=emb.to(DEV))
        wav = wav.cpu()
        if precision == "fp32":
            assert (wav - ref).abs().max() <= 1e-5 * float(ref.abs().max())
        else:
            snr = O.snr_db(ref, wav)
            print("v1 published plan (384 ch) bf16 SNR %.1f dB" % snr)
            assert snr >= 33.0


def test_v1_errors(pkg, synth, cfg):
    h = cfg.tiny_v1_hparams()
    m = make(pkg, h, synth.make_state_dict(h, seed=5), "fp32")
    latent = synth.make_latent(1, 8, h["gpt_dim"]).to(DEV)
    with pytest.raises(RuntimeError):
        m(latent)                                                  # no encoder module and no embedding
    with pytest.raises(RuntimeError):
        m(latent.cpu(), speaker_embedding=torch.zeros(1, h["speaker_embedding_dim"]))
    with pytest.raises(RuntimeError):
        m(latent[:, :, :-1].contiguous(), speaker_embedding=torch.zeros(1, h["speaker_embedding_dim"], device=DEV))
