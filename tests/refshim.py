"""Drive the REAL reference (imported from /root/reference) -- build container only.

Used by tests/golden/make_golden.py to produce the committed fixtures and by the
``live reference`` tests (skipped when /root/reference is absent, e.g. on the GPU box).
Mirrors runner.py:32-45 (build_models) and trainers/ddpm_show_trainer.py:52-80 /
:163-198 (diffusion construction, generate_batch) without mmcv/wandb/lmdb.
"""
import argparse
import os
import sys

REF = "/root/reference"


def have_reference():
    return os.path.isdir(os.path.join(REF, "models"))


def make_opt(cfg, **over):
    """The ~35 Namespace fields the reference models read."""
    opt = argparse.Namespace(
        model_base="transformer_encoder", cond_projection="mlp_includeX", cond_residual=True,
        expCondition_gesture_only=None, gesCondition_expression_only=False, addTextCond=False,
        addEmoCond=False, expAddHubert=False, addHubert=True, addWav2Vec2=False, encode_hubert=True,
        encode_wav2vec2=False, unidiffuser=True, classifier_free=cfg["classifier_free"],
        cond_scale=cfg["cond_scale"], null_cond_prob=0.2, separate=None, ExprID_off=False,
        ExprID_off_uncond=False, no_style=False, visualize_unify_x0_step=None,
        same_overlap_noisy=False, fix_head_var=False, no_repaint=False, no_resample=False,
        addBlend=True, timestep_respacing="ddim25", jump_length=3, jump_n_sample=5,
        overlap_len=0, dataset_name=cfg["dataset_name"], dim_pose=cfg["dim_pose"],
        expression_dim=cfg["expression_dim"], split_pos=cfg["dim_pose"],
        expression_only=False, gesture_only=False, PE="pe_sinu", ddim=True,
        net_dim_pose=cfg["net_dim_pose"], n_poses=cfg["n_poses"], diffusion_steps=1000,
    )
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def build_reference(cfg, sd, opt=None):
    """-> (UniDiffuser in eval mode with ``sd`` loaded strictly, opt)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.transformer import UniDiffuser

    opt = opt or make_opt(cfg)
    model = UniDiffuser(opt=opt, input_feats=cfg["net_dim_pose"], audio_dim=cfg["audio_dim"],
                        aud_latent_dim=cfg["aud_latent_dim"], style_dim=cfg["style_dim"],
                        num_frames=cfg["n_poses"], num_layers=cfg["num_layers"],
                        latent_dim=cfg["latent_dim"], pe_type="pe_sinu")
    model.load_state_dict(sd, strict=True)
    return model.eval(), opt


def build_diffusion(opt, ddim=True, steps=1000):
    """show:52-80: SpacedDiffusion('ddim25') when --ddim else GaussianDiffusion."""
    from models.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                           get_named_beta_schedule)
    from models.respace import SpacedDiffusion, space_timesteps

    betas = get_named_beta_schedule("linear", steps)
    kw = dict(opt=opt, betas=betas, model_mean_type=ModelMeanType.EPSILON,
              model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    if ddim:
        return SpacedDiffusion(use_timesteps=space_timesteps(steps, "ddim25"), rescale_timesteps=False, **kw)
    return GaussianDiffusion(**kw)


def generate_batch(diffusion, model, opt, mel, person_id, hubert, dim_pose, inpaint_dict, ddim=True,
                   noise=None):
    """show:163-198 restated (the trainer itself needs mmcv/wandb)."""
    import torch

    B, T = mel.shape[0], mel.shape[1]
    kw = dict(audio_emb=mel, length=torch.LongTensor([T] * B), person_id=person_id,
              add_cond={"pretrain_aud_feat": hubert}, y=inpaint_dict, pe_type=opt.PE)
    fn = diffusion.ddim_sample_loop if ddim else diffusion.p_sample_loop
    return fn(model, (B, T, dim_pose), noise=noise, clip_denoised=False, progress=False, model_kwargs=kw)
