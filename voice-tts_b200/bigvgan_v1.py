"""Drop-in for the IndexTTS-v1 speaker-conditioned generator, `indextts/BigVGAN/models.py::BigVGAN` (:130-275), as
`indextts/infer.py:112-119` builds it and `:476,646` call it:

    wav, _ = self.bigvgan(latent, auto_conditioning.transpose(1, 2))

Same AMP blocks / Activation1d / kernels as the v2 generator.  What differs
(models.py:212-250): the input is the GPT latent `[B, T, gpt_dim]` (already channels-last), a speaker embedding `e` is
projected by 1x1 convs and added after `conv_pre` (`cond_layer`) and after every ConvTranspose1d (`conds[i]`), the output
always goes through `tanh`, and `forward` returns `(wav, contrastive_loss)`.  On a length-1 sequence a 1x1 conv is a
per-utterance vector, so the native plan folds `cond(e)` into the bias of conv_pre / ups[i] (`bvg_vocoder_fwd_cond`).

The ECAPA-TDNN speaker encoder (`models.py:202`) is NOT part of the hot path: pass the caller's
module as `speaker_encoder=` (its parameters then live under the reference's `speaker_encoder.*` keys) or hand the
embedding in directly with `forward(x, speaker_embedding=e)`.  CUDA (sm_100a) only, no torch fallback.
"""
import torch
import torch.nn as nn

from . import ops
from .bigvgan import BigVGAN as _BigVGANBase
from .config import AttrDict


class BigVGAN(_BigVGANBase):
    def __init__(self, h, use_cuda_kernel: bool = False, precision: str = None, speaker_encoder: nn.Module = None):
        h = AttrDict(dict(h))
        if not h.get("gpt_dim") or not h.get("speaker_embedding_dim"):
            raise ValueError("the v1 generator needs gpt_dim and speaker_embedding_dim in its hyper-parameters")
        if h.get("feat_upsample", False):
            raise NotImplementedError("feat_upsample=True (4x linear interpolation of the latent, models.py:216-222) is not built")
        h["use_tanh_at_final"] = True      # models.py:248
        h["use_bias_at_final"] = True      # models.py:192
        super().__init__(h, use_cuda_kernel=use_cuda_kernel, precision=precision)
        self.cond_in_each_up_layer = bool(h.get("cond_d_vector_in_each_upsampling_layer", False))
        E, c0 = h["speaker_embedding_dim"], h["upsample_initial_channel"]
        self.speaker_encoder = speaker_encoder
        self.cond_layer = nn.Conv1d(E, c0, 1)
        if self.cond_in_each_up_layer:
            self.conds = nn.ModuleList([nn.Conv1d(E, c0 // (2 ** (i + 1)), 1) for i in range(self.num_upsamples)])

    def _native_conditioning(self):
        return 1, self.h["speaker_embedding_dim"], 1 if self.cond_in_each_up_layer else 0

    def forward(self, x, mel_ref=None, lens=None, speaker_embedding=None):
        """x: latent [B, T, gpt_dim] fp32 CUDA.  Returns (wav [B, 1, T*prod(upsample_rates)], None) - the contrastive loss of
        models.py:215-219 is a training-time quantity."""
        if not x.is_cuda:
            raise RuntimeError("BigVGAN (B200 build) runs on CUDA only; got a %s tensor" % x.device)
        if speaker_embedding is None:
            if self.speaker_encoder is None:
                raise RuntimeError("no speaker_encoder module was given: pass speaker_embedding=[B, %d]"
                                   % self.h["speaker_embedding_dim"])
            speaker_embedding = self.speaker_encoder(mel_ref, lens)          # [B, 1, E]  (models.py:213)
        B = x.shape[0]
        e = speaker_embedding.reshape(speaker_embedding.shape[0], -1)[:B].to(device=x.device, dtype=torch.float32).contiguous()
        if self._hid is None or ops._HANDLES[self._hid][3] != x.device.index:
            self._invalidate()
            self._build_native(x.device)
        return ops.vocoder_cond(x.float().contiguous(), e, self._hid), None

    def forward_host(self, *a, **k):
        raise NotImplementedError("the host-buffer entry point is defined for the v2 generator only")
