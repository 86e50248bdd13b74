"""Golden fixtures for the cond_projection / cond_residual variants, produced by the REAL reference (CPU, fp32).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_variants.py

Writes tests/golden/denoise_variants.npz: one denoiser call of the reference UniDiffuser (models/transformer.py:728-770) per
(dataset, cond_projection, cond_residual) with the seeded synthetic weights / inputs of ``diffsheg_b200.synth`` -- key
``{show|beat}_{cond_projection}_{res|nores}`` -> eps, plus the step constants and an input fingerprint per dataset.
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
from make_golden import fingerprint  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = (("show", 2, 88, 12), ("beat", 2, 34, 24))
VARIANTS = [(cp, cr) for cp in ("mlp_includeX", "linear_includeX", "mlp_excludeX", "linear_excludeX") for cr in (True, False)
            if not (cp == "mlp_includeX" and cr)]


def key(name, cp, cr):
    return f"{name}_{cp}_{'res' if cr else 'nores'}"


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    out = {}
    for name, B, T, t_resp in CASES:
        for cp, cr in VARIANTS:
            cfg = synth.make_cfg(name, cond_projection=cp, cond_residual=cr)
            sd = synth.make_state_dict(cfg, seed=1)
            opt = refshim.make_opt(cfg, cond_projection=cp, cond_residual=cr)
            model, opt = refshim.build_reference(cfg, sd, opt)     # strict load: the synthetic key layout IS the reference's
            diff = refshim.build_diffusion(opt, ddim=True)
            inp = synth.make_inputs(cfg, B, T, seed=2)
            t_orig = diff.timestep_map[t_resp]
            a = float(np.float32(diff.sqrt_recip_alphas_cumprod[t_resp]))
            b = float(np.float32(diff.sqrt_recipm1_alphas_cumprod[t_resp]))
            exp_shape = (B, T, cfg["expression_dim"])
            with torch.no_grad():
                eps = model(inp["x_T"], torch.full((B,), t_orig, dtype=torch.long),
                            sqrt_alphas=[torch.full(exp_shape, a), torch.full(exp_shape, b)], audio_emb=inp["mel"],
                            length=torch.LongTensor([T] * B), person_id=inp["person_id"],
                            add_cond={"pretrain_aud_feat": inp["hubert"]}, pe_type="pe_sinu", y={})
            out[key(name, cp, cr)] = eps.numpy()
            out[name + "_consts"] = np.array([B, T, t_resp, t_orig, a, b], dtype=np.float64)
            out[name + "_fp"] = fingerprint(inp["x_T"], inp["mel"], inp["hubert"])
            print(key(name, cp, cr), float(eps.abs().mean()), float(eps.abs().max()))
    np.savez_compressed(os.path.join(OUT, "denoise_variants.npz"), **out)
