"""Generate the committed golden fixtures by running the REAL reference (CPU, fp32).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Inputs are regenerated from seeds by
``diffsheg_b200.synth`` (CPU generators are machine-independent), so each fixture holds
the reference OUTPUT plus a fingerprint of the inputs that produced it.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from diffsheg_b200 import synth  # noqa: E402
import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def fingerprint(*tensors):
    return np.array([float(t.double().sum()) for t in tensors] + [float(t.double().abs().sum()) for t in tensors])


def golden_denoise(name, B, T, t_resp):
    cfg = synth.make_cfg(name)
    sd = synth.make_state_dict(cfg, seed=1)
    model, opt = refshim.build_reference(cfg, sd)
    diff = refshim.build_diffusion(opt, ddim=True)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    x = inp["x_T"]
    t_orig = diff.timestep_map[t_resp]
    a = float(np.float32(diff.sqrt_recip_alphas_cumprod[t_resp]))
    b = float(np.float32(diff.sqrt_recipm1_alphas_cumprod[t_resp]))
    ts = torch.full((B,), t_orig, dtype=torch.long)
    exp_shape = (B, T, cfg["expression_dim"])
    sqrt_alphas = [torch.full(exp_shape, a), torch.full(exp_shape, b)]
    with torch.no_grad():
        eps = model(x, ts, sqrt_alphas=sqrt_alphas, audio_emb=inp["mel"], length=torch.LongTensor([T] * B),
                    person_id=inp["person_id"], add_cond={"pretrain_aud_feat": inp["hubert"]},
                    pe_type="pe_sinu", y={})
    np.savez_compressed(os.path.join(OUT, f"denoise_{name}_B{B}_T{T}_t{t_resp}.npz"), eps=eps.numpy(),
                        t_orig=t_orig, a=a, b=b, fp=fingerprint(x, inp["mel"], inp["hubert"]))
    print("denoise", name, B, T, t_resp, float(eps.abs().mean()))


def golden_loop(name, B, T, overlap, ddim=True, steps=1000, tag="", **opt_over):
    cfg = synth.make_cfg(name)
    sd = synth.make_state_dict(cfg, seed=1)
    opt = refshim.make_opt(cfg, overlap_len=overlap, diffusion_steps=steps, **opt_over)
    model, opt = refshim.build_reference(cfg, sd, opt)
    diff = refshim.build_diffusion(opt, ddim=ddim, steps=steps)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    y = {}
    if overlap > 0:
        g = torch.Generator().manual_seed(5)
        gt = torch.zeros(B, T, cfg["net_dim_pose"])
        gt[:, :overlap] = torch.randn(B, overlap, cfg["net_dim_pose"], generator=g)
        mask = torch.zeros(B, T, cfg["net_dim_pose"], dtype=torch.bool)
        mask[:, :overlap] = True
        y = {"gt": gt, "outpainting_mask": mask}
    torch.manual_seed(1234)  # the loop draws x_T and every per-step noise from the global generator (F11)
    with torch.no_grad():
        out = refshim.generate_batch(diff, model, opt, inp["mel"], inp["person_id"], inp["hubert"],
                                     cfg["net_dim_pose"], y, ddim=ddim)
    fn = f"loop_{name}_B{B}_T{T}_ov{overlap}_{'ddim25' if ddim else 'ddpm%d' % steps}{tag}.npz"
    np.savez_compressed(os.path.join(OUT, fn), sample=out.numpy(), seed=1234,
                        fp=fingerprint(inp["mel"], inp["hubert"]))
    print(fn, float(out.abs().mean()), float(out.abs().max()))


def golden_tables():
    cfg = synth.make_cfg("show")
    opt = refshim.make_opt(cfg)
    refshim.build_reference(cfg, synth.make_state_dict(cfg))  # puts /root/reference on sys.path
    d25 = refshim.build_diffusion(opt, ddim=True)
    d1000 = refshim.build_diffusion(opt, ddim=False)
    from models.scheduler import get_schedule_jump_cjm_ddim, get_schedule_jump_paper
    tabs = {}
    for tag, d in (("d25", d25), ("d1000", d1000)):
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            tabs[f"{tag}_{k}"] = getattr(d, k)
    tabs["d25_timestep_map"] = np.array(d25.timestep_map)
    tabs["jump_25_3_5"] = np.array(get_schedule_jump_cjm_ddim(25, 3, 5))
    tabs["jump_25_3_2"] = np.array(get_schedule_jump_cjm_ddim(25, 3, 2))
    tabs["jump_25_1_1"] = np.array(get_schedule_jump_cjm_ddim(25))
    tabs["jump_paper"] = np.array(get_schedule_jump_paper())
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tabs)
    print("tables", len(tabs))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    golden_tables()
    golden_denoise("show", 2, 88, 12)
    golden_denoise("show", 3, 84, 0)     # ragged last window of a 60 s clip
    golden_denoise("beat", 2, 34, 24)
    golden_denoise("beat", 1, 30, 3)
    golden_loop("show", 1, 88, 0)                       # plain 25 calls
    golden_loop("show", 2, 88, 10)                      # harmonize: 63 calls + 48 undos
    golden_loop("beat", 2, 34, 0)
    golden_loop("beat", 1, 34, 4, jump_n_sample=2, tag="_jn2")   # 27 calls + 12 undos
    golden_loop("beat", 2, 34, 0, ddim=False, steps=40)  # DDPM ancestral path, --diffusion_steps 40
