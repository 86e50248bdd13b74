"""Synthetic configurations, weights and inputs (no reference, no network needed).

The reference ships no checkpoints (SURVEY section 2 "Assets"), so tests and benchmarks use a
seeded synthetic ``state_dict`` with exactly the reference's key/shape layout
(``checkpoint['encoder']``, SURVEY section 8b "Weight contract"; shapes follow
models/transformer.py tr:71-84, tr:100-110, tr:168-176, tr:247-298, tr:349-476, tr:590-699).
Unlike a freshly constructed reference module, NO parameter is left at its zero init
(tr:62-68 zero_module would make every SA/FFN block the identity).
"""
import math

import torch

# dataset-derived constants: runner.py:124-222 of the reference
SHOW = dict(name="show", dataset_name="talkshow", dim_pose=129, expression_dim=103, style_dim=4,
            n_poses=88, fps=30, classifier_free=True, cond_scale=1.25, overlap_len=10)
BEAT = dict(name="beat", dataset_name="beat", dim_pose=141, expression_dim=51, style_dim=30,
            n_poses=34, fps=15, classifier_free=False, cond_scale=1.0, overlap_len=4)
_COMMON = dict(audio_dim=128, hubert_dim=1024, aud_latent_dim=256, latent_dim=512, num_layers=8,
               num_heads=8, ff_size=1024, hubert_enc_dim=128)


def make_cfg(name="show", **overrides):
    base = dict(SHOW if name == "show" else BEAT)
    base.update(_COMMON)
    base.update(overrides)
    base["net_dim_pose"] = base["dim_pose"] + base["expression_dim"]
    return base


def pe_table(d_model=512, period=600, max_seq_len=600):
    """PeriodicPositionalEncoding buffer, tr:19-31 (pe_sinu: period 600 -> [1,1200,d])."""
    pe = torch.zeros(period, d_model)
    position = torch.arange(0, period, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).repeat(1, (max_seq_len // period) + 1, 1)


def state_dict_shapes(cfg):
    """Ordered {key: shape} of the reference UniDiffuser state_dict for ``cfg``."""
    D, E, F_, A = cfg["latent_dim"], cfg["latent_dim"] * 4, cfg["ff_size"], cfg["audio_dim"]
    AL, HE = cfg["aud_latent_dim"], cfg["hubert_enc_dim"]
    shapes = {}

    def lin(prefix, n_out, n_in):
        shapes[prefix + ".weight"] = (n_out, n_in)
        shapes[prefix + ".bias"] = (n_out,)

    def ln(prefix, n):
        shapes[prefix + ".weight"] = (n,)
        shapes[prefix + ".bias"] = (n,)

    def styl(prefix, d):
        lin(prefix + ".emb_layers.1", 2 * d, E)
        ln(prefix + ".norm", d)
        lin(prefix + ".out_layers.2", d, d)

    cond_projection = cfg.get("cond_projection", "mlp_includeX")

    def layer(prefix, d, pre_proj):
        if pre_proj and cond_projection.startswith("linear"):   # tr:281-282
            lin(prefix + ".feat_proj", d, pre_proj)
        elif pre_proj:
            ln(prefix + ".feat_proj.0", pre_proj)
            lin(prefix + ".feat_proj.1", 2 * d, pre_proj)
            lin(prefix + ".feat_proj.3", d, 2 * d)
        ln(prefix + ".sa_block.norm", d)
        for n in ("query", "key", "value"):
            lin(prefix + ".sa_block." + n, d, d)
        styl(prefix + ".sa_block.proj_out", d)
        lin(prefix + ".ffn.linear1", F_, d)
        lin(prefix + ".ffn.linear2", d, F_)
        styl(prefix + ".ffn.proj_out", d)

    lin("time_embed.0", E, D)
    lin("time_embed.2", E, E)
    layer("encoder_aud", A, 0)
    for net, feats, extra in (("encoder_exp", cfg["expression_dim"], 0),
                              ("encoder_ges", cfg["dim_pose"], cfg["expression_dim"])):
        P = (D if cond_projection.endswith("includeX") else 0) + AL + extra + HE   # tr:260-276
        if cfg["classifier_free"]:
            shapes[net + ".null_cond_emb"] = (1, P)
        shapes[net + ".PE.pe"] = (1, 1200, D)
        lin(net + ".joint_embed", D, feats)
        lin(net + ".audio_proj", AL, 2 * A)
        shapes[net + ".hubert_encoder.0.weight"] = (HE, cfg["hubert_dim"], 3)
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[net + ".hubert_encoder.1." + k] = (HE,)
        shapes[net + ".hubert_encoder.1.num_batches_tracked"] = ()
        shapes[net + ".hubert_encoder.3.weight"] = (HE, HE, 3)
        lin(net + ".time_embed.0", E, D)
        lin(net + ".time_embed.2", E, E)
        lin(net + ".pid_embed.0", E, cfg["style_dim"])
        lin(net + ".pid_embed.2", E, E)
        for i in range(cfg["num_layers"]):
            layer(f"{net}.temporal_decoder_blocks.{i}", D, P)
        lin(net + ".out", feats, D)
    return shapes


def make_state_dict(cfg, seed=1):
    """Seeded synthetic weights (CPU fp32), every tensor non-trivial."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in state_dict_shapes(cfg).items():
        if key.endswith("PE.pe"):
            sd[key] = pe_table(cfg["latent_dim"])
        elif key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("running_mean"):
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("null_cond_emb"):
            sd[key] = torch.randn(shape, generator=g)
        elif ".norm." in key or ".feat_proj.0." in key or ".hubert_encoder.1." in key:
            base = 1.0 if key.endswith("weight") else 0.0
            sd[key] = base + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith(".bias"):
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        else:  # linear / conv weight: unit-gain fan-in scaling
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 0.5 if key.endswith("out.weight") else 1.0
            sd[key] = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
    return sd


def make_inputs(cfg, B, T=None, seed=2):
    """Synthetic conditioning + noise for one window: CPU fp32 tensors."""
    T = T or cfg["n_poses"]
    g = torch.Generator().manual_seed(seed)
    mel = torch.randn(B, T, cfg["audio_dim"], generator=g)
    hubert = torch.randn(B, T, cfg["hubert_dim"], generator=g)
    pid = torch.zeros(B, cfg["style_dim"])
    pid[torch.arange(B), torch.arange(B) % cfg["style_dim"]] = 1.0
    g3 = torch.Generator().manual_seed(seed + 1)
    x_T = torch.randn(B, T, cfg["net_dim_pose"], generator=g3)
    return dict(mel=mel, hubert=hubert, person_id=pid, x_T=x_T)


def make_opt(cfg, **over):
    """The sampler-relevant subset of the reference's ``opt`` Namespace (options/base_options.py:16-128,
    runner.py:124-222) for synthetic runs: what FusedGaussianDiffusion / generate_batch read."""
    import argparse

    opt = argparse.Namespace(
        dataset_name=cfg["dataset_name"], dim_pose=cfg["dim_pose"], expression_dim=cfg["expression_dim"],
        split_pos=cfg["dim_pose"], net_dim_pose=cfg["net_dim_pose"], n_poses=cfg["n_poses"],
        style_dim=cfg["style_dim"], classifier_free=cfg["classifier_free"], cond_scale=cfg["cond_scale"],
        ddim=True, timestep_respacing="ddim25", diffusion_steps=1000, overlap_len=0, addBlend=True,
        jump_length=3, jump_n_sample=5, no_resample=False, no_repaint=False, same_overlap_noisy=False,
        fix_head_var=False, PE="pe_sinu", model_mean_type="epsilon", unidiffuser=True)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt
