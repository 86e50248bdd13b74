"""Functional restatement of the reference denoiser (test infrastructure only).

Every function takes the reference ``state_dict`` (``checkpoint['encoder']`` key
layout) and plain tensors; dtype/device follow the inputs, so the
same code gives the fp32 CPU baseline, an fp64 "truth" run and (on a GPU box) an
eager torch-cuda run of the reference op stream.

Restated: unidiffuser=True, model_base=transformer_encoder, addHubert=encode_hubert=True,
PE=pe_sinu, eps prediction, with every cond_projection the reference's UniDiffuser can
run (mlp_includeX -- the shipped one --, linear_includeX, mlp_excludeX, linear_excludeX;
tr:262-263,281-289,302-338) and cond_residual on or off.  'none' builds
(tr:641-650) but raises NotImplementedError in every MotionTransformer layer (tr:323-324).
"""
import math

import torch
import torch.nn.functional as F

COND_PROJECTIONS = ("mlp_includeX", "linear_includeX", "mlp_excludeX", "linear_excludeX")


def timestep_embedding(timesteps, dim, max_period=10000):
    """tr:42-59 -- [cos | sin] (cos first) of t * max_period^(-i/half)."""
    half = dim // 2
    freqs = torch.exp(
        -math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half
    ).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def positional_encoding(T, d_model, dtype, device, period=600):
    """tr:19-38 with period=600 (pe_sinu): rows [0,T) of the plain sinusoid table."""
    assert T <= period
    pe = torch.zeros(T, d_model)
    position = torch.arange(0, T, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(device=device, dtype=dtype)


class _SD:
    """state_dict view with a key prefix, cast to the compute dtype/device."""

    def __init__(self, sd, prefix, dtype, device):
        self.sd, self.prefix, self.dtype, self.device = sd, prefix, dtype, device

    def __call__(self, key):
        return self.sd[self.prefix + key].to(device=self.device, dtype=self.dtype)

    def sub(self, prefix):
        return _SD(self.sd, self.prefix + prefix, self.dtype, self.device)


def _linear(p, name, x):
    return F.linear(x, p(name + ".weight"), p(name + ".bias"))


def _layer_norm(p, name, x):
    w = p(name + ".weight")
    return F.layer_norm(x, (w.shape[0],), w, p(name + ".bias"), 1e-5)


def stylization(p, h, emb):
    """StylizationBlock.forward tr:86-97."""
    emb_out = _linear(p, "emb_layers.1", F.silu(emb)).unsqueeze(1)
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = _layer_norm(p, "norm", h) * (1 + scale) + shift
    return _linear(p, "out_layers.2", F.silu(h))


def linear_self_attention(p, x, emb, num_head):
    """LinearTemporalSelfAttention.forward tr:112-130 (src_mask is all ones, F8)."""
    B, T, D = x.shape
    H = num_head
    xn = _layer_norm(p, "norm", x)
    query = _linear(p, "query", xn)
    key = _linear(p, "key", xn)
    query = F.softmax(query.view(B, T, H, -1), dim=-1)
    key = F.softmax(key.view(B, T, H, -1), dim=1)
    value = _linear(p, "value", xn).view(B, T, H, -1)
    attention = torch.einsum("bnhd,bnhl->bhdl", key, value)
    y = torch.einsum("bnhd,bhdl->bnhl", query, attention).reshape(B, T, D)
    return x + stylization(p.sub("proj_out."), y, emb)


def ffn(p, x, emb):
    """FFN.forward tr:178-181 (GELU = erf form)."""
    y = _linear(p, "linear2", F.gelu(_linear(p, "linear1", x)))
    return x + stylization(p.sub("proj_out."), y, emb)


def transformer_layer(p, x, xf, emb, add_cond, null_cond_emb, num_head, cfg_double,
                      cond_projection="mlp_includeX", cond_residual=True):
    """LinearTemporalDiffusionTransformerLayer.forward tr:300-346, eval mode.

    cond_projection (tr:304-324): *_includeX projects cat(x, xf, add_cond), *_excludeX projects cat(xf, add_cond) only;
    mlp_* = LayerNorm -> Linear -> SiLU -> Linear, linear_* = one Linear (tr:281-289).
    The input is added back when cond_residual is set OR the projection excludes x (tr:302-303,337-338)."""
    assert cond_projection in COND_PROJECTIONS, cond_projection
    residual = cond_residual or cond_projection.endswith("excludeX")
    x_ori = x
    if xf is not None:
        parts = ([x] if cond_projection.endswith("includeX") else []) + [xf] + ([] if add_cond is None else [add_cond])
        x = torch.cat(parts, -1)
        if cfg_double:
            # tr:330-332: rows [0, B'/2) take the learned null row for the WHOLE feat_proj input.
            n = x.shape[0]
            mask = (torch.linspace(0, 1, n) < 0.5).to(x.device)
            null = null_cond_emb.repeat(x.shape[1], 1).unsqueeze(0)
            x = torch.where(mask.unsqueeze(1).unsqueeze(2), null, x)
        if cond_projection.startswith("mlp"):
            fp = p.sub("feat_proj.")
            x = _layer_norm(fp, "0", x)
            x = _linear(fp, "3", F.silu(_linear(fp, "1", x)))
        else:
            x = _linear(p, "feat_proj", x)
    if residual:
        x = x + x_ori  # NB with xf=None (encoder_aud) this doubles x (tr:337-338)
    x = linear_self_attention(p.sub("sa_block."), x, emb, num_head)
    return ffn(p.sub("ffn."), x, emb)


def hubert_encoder(p, feat):
    """tr:436-442 applied as tr:515: Conv1d(1024,128,3,p1) -> BN(eval) -> GELU -> Conv1d(128,128,3,p1)."""
    z = feat.transpose(-1, -2)
    z = F.conv1d(z, p("0.weight"), None, padding=1)
    z = F.batch_norm(z, p("1.running_mean"), p("1.running_var"), p("1.weight"), p("1.bias"),
                     training=False, eps=1e-5)
    z = F.gelu(z)
    z = F.conv1d(z, p("3.weight"), None, padding=1)
    return z.transpose(-1, -2)


def motion_transformer(p, x, timesteps, audio_emb, person_id, hubert, exp_cond,
                       num_layers, num_head, latent_dim, cond_scale, classifier_free,
                       cond_projection="mlp_includeX", cond_residual=True):
    """MotionTransformer.forward tr:496-587.

    audio_emb: [B,T,256] (mel | audio_feat); hubert: raw [B,T,1024];
    exp_cond: None (expression net) or [B,T,expression_dim] (gesture net, tr:506-507).
    """
    add_cond = hubert_encoder(p.sub("hubert_encoder."), hubert)
    if exp_cond is not None:
        add_cond = torch.cat((add_cond, exp_cond), dim=-1)
    cfg_double = bool(classifier_free and cond_scale != 1)
    if cfg_double:  # tr:537-544
        x = torch.cat([x] * 2)
        timesteps = torch.cat([timesteps] * 2)
        audio_emb = torch.cat([audio_emb] * 2)
        person_id = torch.cat([person_id] * 2)
        add_cond = torch.cat([add_cond] * 2)
    temb = timestep_embedding(timesteps, latent_dim).to(x.dtype)
    emb = _linear(p, "time_embed.2", F.silu(_linear(p, "time_embed.0", temb))) + \
        _linear(p, "pid_embed.2", F.silu(_linear(p, "pid_embed.0", person_id.to(x.dtype))))
    B, T = x.shape[0], x.shape[1]
    h = _linear(p, "joint_embed", x)
    h = h + positional_encoding(T, latent_dim, x.dtype, x.device)[None]
    xf = _linear(p, "audio_proj", audio_emb)
    null = p("null_cond_emb") if classifier_free else None
    for i in range(num_layers):
        h = transformer_layer(p.sub(f"temporal_decoder_blocks.{i}."), h, xf, emb, add_cond,
                              null, num_head, cfg_double, cond_projection, cond_residual)
    out = _linear(p, "out", h).view(B, T, -1)
    if cfg_double:  # tr:585-586
        half = out.shape[0] // 2
        out = out[:half] + cond_scale * (out[half:] - out[:half])
    return out


def unidiffuser_forward(sd, cfg, x, timesteps, sqrt_alphas, audio_emb, person_id, hubert,
                        dtype=torch.float32):
    """UniDiffuser.forward tr:728-770.

    x [B,T,dim_pose+expression_dim] (gesture first), timesteps [B] (ORIGINAL 0..999
    timesteps, i.e. after rs:119-124), sqrt_alphas = (sqrt_recip_alphas_cumprod_t,
    sqrt_recipm1_alphas_cumprod_t) broadcastable to the expression block (gd:527-532),
    audio_emb = mel [B,T,128], person_id [B,style_dim], hubert [B,T,1024].
    cfg: dict with dim_pose, expression_dim, num_layers, num_heads, latent_dim,
    classifier_free, cond_scale and optionally cond_projection / cond_residual
    (defaults: the shipped mlp_includeX / True).
    """
    cp, cr = cfg.get("cond_projection", "mlp_includeX"), bool(cfg.get("cond_residual", True))
    device = x.device
    p = _SD(sd, "", dtype, device)
    x = x.to(dtype)
    audio_emb = audio_emb.to(dtype)
    hubert = hubert.to(dtype)
    L, H, D = cfg["num_layers"], cfg["num_heads"], cfg["latent_dim"]
    temb = timestep_embedding(timesteps, D).to(dtype)
    emb = _linear(p, "time_embed.2", F.silu(_linear(p, "time_embed.0", temb)))
    audio_feat = transformer_layer(p.sub("encoder_aud."), audio_emb, None, emb, None, None, H, False, cp, cr)
    audio_emb = torch.cat((audio_emb, audio_feat), dim=-1)
    gesture, expression = torch.split(x, [cfg["dim_pose"], cfg["expression_dim"]], dim=-1)
    exp_noise = motion_transformer(p.sub("encoder_exp."), expression, timesteps, audio_emb, person_id,
                                   hubert, None, L, H, D, cfg["cond_scale"], cfg["classifier_free"], cp, cr)
    a, b = sqrt_alphas
    expr_cond = a * expression - b * exp_noise  # tr:717-725, tr:749
    ges_noise = motion_transformer(p.sub("encoder_ges."), gesture, timesteps, audio_emb, person_id,
                                   hubert, expr_cond, L, H, D, cfg["cond_scale"], cfg["classifier_free"], cp, cr)
    return torch.cat((ges_noise, exp_noise), dim=-1)
