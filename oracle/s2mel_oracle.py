"""ORACLE (test infrastructure; never imported by the product) - CPU restatement of the s2mel tail.

Restates, in plain torch CPU ops on FOLDED weights (native tensor names, see voice-tts_b200/synth.s2mel_tail_spec):
  * DiT.forward after the transformer          indextts/s2mel/modules/diffusion_transformer.py:245-256
  * TimestepEmbedder                           diffusion_transformer.py:41-57
  * FinalLayer                                 diffusion_transformer.py:95-99
  * WN.forward                                 indextts/s2mel/modules/wavenet.py:140-164
  * SConv1d reflect padding                    indextts/s2mel/modules/encodec.py:96-113,212-228
  * fused_add_tanh_sigmoid_multiply            indextts/s2mel/modules/commons.py:133-139
  * BASECFM.solve_euler                        indextts/s2mel/modules/flow_matching.py:57-113
Pinned: tests/golden/s2mel_tail.npz is produced by oracle/make_golden.py s2mel from the UNMODIFIED reference DiT.forward /
BASECFM.solve_euler (tests/test_oracle.py::test_s2mel_tail_oracle_matches_reference_goldens).
"""
import torch
import torch.nn.functional as F


def timestep_embedder(sd, prefix, t, dtype=torch.float32):
    """diffusion_transformer.py:41-57: args = 1000 * t * freqs; [cos | sin]; Linear - SiLU - Linear"""
    args = 1000 * t[:, None].float() * sd[prefix + "freqs"][None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(dtype)
    h = F.linear(emb, sd[prefix + "mlp.0.weight"].to(dtype), sd[prefix + "mlp.0.bias"].to(dtype))
    return F.linear(F.silu(h), sd[prefix + "mlp.2.weight"].to(dtype), sd[prefix + "mlp.2.bias"].to(dtype))


def sconv1d(x, w, b):
    """encodec.SConv1d, non-causal, stride 1, dilation 1: reflect pad (k-1) split right = total // 2, left = rest"""
    k = w.shape[-1]
    total = k - 1
    right = total // 2
    left = total - right
    if total:
        x = F.pad(x, (left, right), mode="reflect")
    return F.conv1d(x, w, b)


def wn_forward(sd, cfg, x, x_mask, g, dtype=torch.float32):
    """wavenet.py:140-164; x [B, H, T], x_mask [B, 1, T], g [B, H, 1]"""
    H, L = cfg["hidden"], cfg["n_layers"]
    W = lambda k: sd[k].to(dtype)
    output = torch.zeros_like(x)
    g = sconv1d(g, W("wavenet.cond_layer.weight"), W("wavenet.cond_layer.bias"))
    for i in range(L):
        x_in = sconv1d(x, W("wavenet.in_layers.%d.weight" % i), W("wavenet.in_layers.%d.bias" % i))
        g_l = g[:, i * 2 * H:(i + 1) * 2 * H, :]
        in_act = x_in + g_l
        acts = torch.tanh(in_act[:, :H, :]) * torch.sigmoid(in_act[:, H:, :])
        rs = sconv1d(acts, W("wavenet.res_skip_layers.%d.weight" % i), W("wavenet.res_skip_layers.%d.bias" % i))
        if i < L - 1:
            x = (x + rs[:, :H, :]) * x_mask
            output = output + rs[:, H:, :]
        else:
            output = output + rs
    return output * x_mask


def tail_forward(sd, cfg, x_res, x_lens, t, t1, dtype=torch.float32):
    """diffusion_transformer.py:245-256.  x_res [B, T, D], x_lens [B] or None, t [B], t1 [B, H] -> [B, C, T]"""
    W = lambda k: sd[k].to(dtype)
    x_res, t1 = x_res.to(dtype), t1.to(dtype)
    B, T, _ = x_res.shape
    H = cfg["hidden"]
    if x_lens is None:
        x_mask = torch.ones(B, 1, T, dtype=dtype)
    else:
        x_mask = (torch.arange(T)[None, :] < x_lens.long()[:, None]).unsqueeze(1).to(dtype)   # commons.sequence_mask
    x = F.linear(x_res, W("conv1.weight"), W("conv1.bias")).transpose(1, 2)
    t2 = timestep_embedder(sd, "t_embedder2.", t, dtype)
    x = wn_forward(sd, cfg, x, x_mask, t2.unsqueeze(2), dtype).transpose(1, 2) + F.linear(x_res, W("res_projection.weight"),
                                                                                         W("res_projection.bias"))
    mod = F.linear(F.silu(t1), W("final_layer.adaLN_modulation.1.weight"), W("final_layer.adaLN_modulation.1.bias"))
    shift, scale = mod.chunk(2, dim=1)
    x = F.layer_norm(x, (H,), None, None, 1e-6) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
    x = F.linear(x, W("final_layer.linear.weight"), W("final_layer.linear.bias")).transpose(1, 2)
    return F.conv1d(x, W("conv2.weight"), W("conv2.bias"))


def solve_euler(estimator, x, x_lens, prompt, mu, style, t_span, inference_cfg_rate):
    """flow_matching.py:57-113 (zero_prompt_speech_token False), fp32 tensor ops in the reference's order"""
    t = t_span[0]
    prompt_len = prompt.size(-1)
    prompt_x = torch.zeros_like(x)
    prompt_x[..., :prompt_len] = prompt[..., :prompt_len]
    x = x.clone()
    x[..., :prompt_len] = 0
    for step in range(1, len(t_span)):
        dt = t_span[step] - t_span[step - 1]
        if inference_cfg_rate > 0:
            stacked = estimator(torch.cat([x, x], dim=0), torch.cat([prompt_x, torch.zeros_like(prompt_x)], dim=0), x_lens,
                                torch.cat([t.unsqueeze(0), t.unsqueeze(0)], dim=0),
                                torch.cat([style, torch.zeros_like(style)], dim=0), torch.cat([mu, torch.zeros_like(mu)], dim=0))
            dphi_dt, cfg_dphi_dt = stacked.chunk(2, dim=0)
            dphi_dt = (1.0 + inference_cfg_rate) * dphi_dt - inference_cfg_rate * cfg_dphi_dt
        else:
            dphi_dt = estimator(x, prompt_x, x_lens, t.unsqueeze(0), style, mu)
        x = x + dt * dphi_dt
        t = t + dt
        x[:, :, :prompt_len] = 0
    return x


def toy_estimator(x, prompt_x, x_lens, t, style, mu):
    """A stand-in estimator whose arithmetic is exact in fp32 on any device (powers of two, flips, adds of small dyadic
    numbers are not needed to be exact - IEEE adds/muls round identically on CPU and GPU): used by the solver parity tests.
    `t` is the reference's stacked time, shape [2] with CFG and [1] without WHATEVER the batch is (flow_matching.py:93)."""
    return 0.5 * x.flip(-1) - 0.25 * prompt_x + 0.125 * mu.transpose(1, 2)[:, :x.shape[1], :] + (t.reshape(-1)[0] * 0.5) \
        + style[:, :1, None] * 0.25
