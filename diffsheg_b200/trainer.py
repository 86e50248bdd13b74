"""SEAM #3: the trainers' ``generate_batch`` and the overlapped-window loop, over the fused path.

* ``generate_batch`` mirrors trainers/ddpm_show_trainer.py:163-198 / ddpm_beat_trainer.py:185-220.
* ``patch_trainer`` swaps a live reference trainer's ``diffusion`` / ``diffusion_ddim_val`` for the
  fused classes so runner.py and its --ddim / --timestep_respacing / --cond_scale / --overlap_len /
  --jump_* flags keep working unchanged.
* ``generate_long`` is the window loop of ``test_arbitrary_len`` (show:801-906): windows of
  ``n_poses`` with stride ``n_poses - overlap_len``; window ii > 0 repaints its first
  ``overlap_len`` frames from the tail of window ii-1.  It stays in Python (north_star).
"""
import torch

from .diffusion import FusedGaussianDiffusion, FusedSpacedDiffusion, get_named_beta_schedule, space_timesteps


def build_diffusions(opt, precision="bf16", max_batch=None):
    """show:52-80: (GaussianDiffusion over all steps, SpacedDiffusion('ddim25') when opt.ddim)."""
    steps = getattr(opt, "diffusion_steps", 1000)
    betas = get_named_beta_schedule("linear", steps)
    kw = dict(opt=opt, betas=betas, model_mean_type="EPSILON", model_var_type="FIXED_SMALL",
              precision=precision, max_batch=max_batch)
    full = FusedGaussianDiffusion(**kw)
    ddim = None
    if getattr(opt, "ddim", False):
        # NB the step count is hard-coded to 'ddim25' in the reference (show:73), not --timestep_respacing
        ddim = FusedSpacedDiffusion(use_timesteps=space_timesteps(steps, "ddim25"), rescale_timesteps=False, **kw)
    return full, ddim


def patch_trainer(trainer, precision="bf16", max_batch=None):
    """Swap the sampler objects of a reference DDPMTrainer_show / DDPMTrainer_beat in place."""
    full, ddim = build_diffusions(trainer.opt, precision=precision, max_batch=max_batch)
    # the trainer also TRAINS through self.diffusion (training_losses, show:132; q_sample; the schedule sampler built on it,
    # show:53): the fused object answers the sampling entry points and hands every other attribute to the original
    full._fallback = getattr(trainer, "diffusion", None)
    trainer.diffusion = full
    if ddim is not None:
        ddim._fallback = getattr(trainer, "diffusion_ddim_val", None)
        trainer.diffusion_ddim_val = ddim
    return trainer


def generate_batch(opt, encoder, diffusion, audio_emb, p_id, dim_pose, add_cond=None, inpaint_dict=None, noise=None):
    """show:163-198.  ``diffusion`` is the DDIM object when opt.ddim else the full-chain one."""
    B, T = audio_emb.shape[0], audio_emb.shape[1]
    cur_len = torch.full((B,), T, dtype=torch.long)
    model_kwargs = {"audio_emb": audio_emb, "length": cur_len, "person_id": p_id, "add_cond": add_cond or {},
                    "y": inpaint_dict if inpaint_dict is not None else {}, "pe_type": getattr(opt, "PE", "pe_sinu")}
    if getattr(opt, "ddim", False):
        return diffusion.ddim_sample_loop(encoder, (B, T, dim_pose), noise=noise, clip_denoised=False, progress=False,
                                          model_kwargs=model_kwargs)
    return diffusion.p_sample_loop(encoder, (B, T, dim_pose), noise=noise, clip_denoised=False, progress=False,
                                   model_kwargs=model_kwargs)


def get_windows(x, size, step):
    """show:801-819 for a tensor [B, frames, C] (or a dict of them)."""
    if isinstance(x, dict):
        per_key = {k: get_windows(v, size, step) for k, v in x.items()}
        n = len(next(iter(per_key.values())))
        return [{k: per_key[k][i] for k in per_key} for i in range(n)]
    frames = x.shape[1]
    if frames <= size:
        return [x]
    win_num = (frames - (size - step)) / float(step)
    out = [x[:, i * step:i * step + size, ...] for i in range(int(win_num))]
    if win_num - int(win_num) != 0:
        out.append(x[:, int(win_num) * step:, ...])
    return out


def generate_long(opt, encoder, diffusion, audio_emb, p_id, dim_pose, add_cond, motions=None):
    """show:864-906: sequential windows of one (batch of) clip(s); returns [B, frames, dim_pose] on the GPU.

    Unlike the reference there is no per-window D2H copy (show:897): windows are concatenated on
    the device and the caller copies once.  ``motions`` ([B, frames, dim_pose]) is only read for
    ``opt.fix_very_first`` (show:885-888: the first window repaints its head from the tail of its own
    ground-truth window).
    """
    n_poses, ov = opt.n_poses, opt.overlap_len
    step = n_poses - ov
    fix_first = bool(getattr(opt, "fix_very_first", False)) and ov > 0
    if fix_first and motions is None:
        raise ValueError("opt.fix_very_first needs the ground-truth motions (show:885-888)")
    audio_list = get_windows(audio_emb, n_poses, step)
    cond_list = get_windows(add_cond, n_poses, step) if add_cond else [add_cond or {}] * len(audio_list)
    outs, prev, prev_tail = [], None, None
    same_noisy = bool(getattr(opt, "same_overlap_noisy", False)) and ov > 0 and bool(getattr(opt, "ddim", False))
    for ii, (aud, cond) in enumerate(zip(audio_list, cond_list)):
        inpaint = {}
        if ov > 0:
            shape = (aud.shape[0], aud.shape[1], dim_pose)
            gt = torch.zeros(shape, device=aud.device)
            mask = torch.zeros(shape, dtype=torch.bool, device=aud.device)
            if ii > 0:
                mask[:, :ov, :] = True
                gt[:, :ov, :] = prev[:, -ov:, :].to(aud.device)
            elif fix_first:
                mask[:, :ov, :] = True
                gt[:, :ov, :] = get_windows(motions, n_poses, step)[0][:, -ov:, :].to(aud.device)
            inpaint = {"gt": gt, "outpainting_mask": mask, "clip_idx": ii}   # beat:1006
            if ii > 0 and same_noisy:
                inpaint["previous_noisy_tail"] = prev_tail                      # beat:1022-1023
        prev = generate_batch(opt, encoder, diffusion, aud, p_id, dim_pose, cond, inpaint)
        if same_noisy:                                                          # beat:1026-1028
            prev, prev_tail = prev["sample"], prev["saved_noisy_tail"]
        outs.append(prev if ii == len(audio_list) - 1 else prev[:, :step])
    return torch.cat(outs, dim=1)
