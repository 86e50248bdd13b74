"""All five BASELINE.json configurations on one GPU (device-timed, synthetic inputs); prints one JSON line each."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsheg_b200 import (FusedGaussianDiffusion, FusedSpacedDiffusion, FusedUniDiffuser, generate_batch, generate_long,  # noqa: E402
                           get_named_beta_schedule, space_timesteps, synth)


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps, out


def setup(name, B, T, ddim=True, steps=1000, **opt_over):
    cfg = synth.make_cfg(name)
    sd = synth.make_state_dict(cfg, seed=1)
    T = T or cfg["n_poses"]
    eng = FusedUniDiffuser(sd, cfg, precision="bf16", max_batch=B, max_frames=min(T, cfg["n_poses"]))
    opt = synth.make_opt(cfg, ddim=ddim, diffusion_steps=steps, **opt_over)
    betas = get_named_beta_schedule("linear", steps)
    diff = FusedSpacedDiffusion(space_timesteps(steps, "ddim25"), opt=opt, betas=betas) if ddim else FusedGaussianDiffusion(opt=opt, betas=betas)
    inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, B, T, seed=2).items()}
    return cfg, eng, opt, diff, inp


def emit(tag, frames, ms, wall, extra=None):
    print(json.dumps(dict(config=tag, frames=frames, ms_device=round(ms, 3), ms_wall=round(wall, 3),
                          frames_per_s=round(frames / (ms / 1e3), 1), **(extra or {}))), flush=True)


def run_batch(tag, name, B, ddim=True, steps=1000, overlap=0, reps=1):
    cfg, eng, opt, diff, inp = setup(name, B, None, ddim, steps, overlap_len=overlap)
    T, Dm = cfg["n_poses"], cfg["net_dim_pose"]
    y = {}
    if overlap:
        y = {"gt": torch.randn(B, T, Dm, device="cuda"), "outpainting_mask": torch.zeros(B, T, Dm, dtype=torch.bool, device="cuda")}
        y["outpainting_mask"][:, :overlap] = True
    fn = lambda: generate_batch(opt, eng, diff, inp["mel"], inp["person_id"], Dm, {"pretrain_aud_feat": inp["hubert"]}, y)
    ms, wall, out = timed(fn, reps)
    assert torch.isfinite(out).all()
    emit(tag, B * T, ms, wall, dict(stats=diff.last_stats))


which = sys.argv[1:] or ["1", "3", "4", "5"]
if "1" in which:
    run_batch("1a: BEAT T=34 B=1 ddim25 overlap 0", "beat", 1, reps=5)
    run_batch("1b: BEAT T=34 B=1 ddim25 overlap 4 (63 calls + 48 undo)", "beat", 1, overlap=4, reps=3)
    run_batch("1c: SHOW T=88 B=1 ddim25 CFG overlap 0", "show", 1, reps=5)
if "3" in which:
    run_batch("3: BEAT T=34 B=2500 ddpm1000 (p_sample_loop)", "beat", 2500, ddim=False)
if "4" in which:
    cfg, eng, opt, diff, _ = setup("show", 8, 88, overlap_len=10)
    frames = 1800
    inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, 8, frames, seed=3).items()}
    fn = lambda: generate_long(opt, eng, diff, inp["mel"], inp["person_id"], cfg["net_dim_pose"], {"pretrain_aud_feat": inp["hubert"]})
    ms, wall, out = timed(fn)
    assert out.shape == (8, frames, cfg["net_dim_pose"]) and torch.isfinite(out).all()
    emit("4: SHOW 60 s clips (1800 frames), overlap 10, 8 clips on one GPU, 23 windows each", 8 * frames, ms, wall)
if "5" in which:
    run_batch("5: SHOW T=88 B=4096 ddim25 CFG (one GPU)", "show", 4096)
