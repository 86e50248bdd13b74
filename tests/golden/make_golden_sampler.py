"""Fixtures for the SAMPLER control flow that the network-driven goldens do not reach (build container only: imports the real
reference from /root/reference).  The denoiser is a cheap closed-form toy eps(x, t) -- the sampler code under test does not care
what produces eps -- so the 2 410-call DDPM RePaint loop (gd:843-920, sch:150-176) and a 3-window --same_overlap_noisy chain
(gd:1040-1042, :1058-1060, :1155-1159; beat:1006-1028) run in seconds on a CPU.

    python tests/golden/make_golden_sampler.py      ->  tests/golden/sampler_*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402
from diffsheg_b200 import synth  # noqa: E402

B, T, D, OV = 2, 12, 6, 3


def toy_eps(x, t_orig):
    """eps(x, t): bounded, mixes channels and frames, depends on the ORIGINAL timestep (what _WrappedModel passes, rs:119-124)."""
    return 0.9 * torch.tanh(0.7 * x.flip(-1) + 0.3 * x.roll(1, 1) + 0.002 * float(t_orig))


class ToyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))   # `next(model.parameters()).device` (gd:1181)
        self.calls = []

    def forward(self, x, ts, **kw):
        assert bool((ts == ts[0]).all())
        self.calls.append(int(ts[0]))
        return toy_eps(x, int(ts[0]))


def inpaint(seed=5):
    g = torch.Generator().manual_seed(seed)
    gt = torch.zeros(B, T, D)
    gt[:, :OV] = torch.randn(B, OV, D, generator=g)
    mask = torch.zeros(B, T, D, dtype=torch.bool)
    mask[:, :OV] = True
    return gt, mask


def main():
    sys.path.insert(0, refshim.REF)
    cfg = dict(synth.make_cfg("beat"), dim_pose=4, expression_dim=2, net_dim_pose=D, n_poses=T)
    model = ToyModel().eval()

    # ---- (A) DDPM RePaint loop: p_sample_loop with a True mask -> get_schedule_jump_paper (t_T = 250, jump 10 x 10)
    opt = refshim.make_opt(cfg, ddim=False, overlap_len=OV, split_pos=4)
    diff = refshim.build_diffusion(opt, ddim=False, steps=1000)
    gt, mask = inpaint()
    torch.manual_seed(11)
    kw = dict(y={"gt": gt.clone(), "outpainting_mask": mask}, pe_type="pe_sinu")
    with torch.no_grad():
        out = diff.p_sample_loop(model, (B, T, D), clip_denoised=False, progress=False, model_kwargs=kw)
    np.savez(os.path.join(HERE, "sampler_ddpm_harmonize_toy.npz"), seed=11, gt=gt.numpy(), sample=out.numpy(), calls=len(model.calls),
             first_t=model.calls[0], last_t=model.calls[-1])
    print("ddpm harmonize: calls", len(model.calls), "first/last t", model.calls[0], model.calls[-1], "absmax", float(out.abs().max()))

    # ---- (B) --same_overlap_noisy chain of 3 windows (window ii > 0 repaints its head from window ii-1's saved noisy tails)
    model.calls.clear()
    opt = refshim.make_opt(cfg, ddim=True, overlap_len=OV, split_pos=4, same_overlap_noisy=True)
    diff = refshim.build_diffusion(opt, ddim=True, steps=1000)
    torch.manual_seed(23)
    samples, prev, tail = [], None, None
    with torch.no_grad():
        for ii in range(3):
            gtw = torch.zeros(B, T, D)
            maskw = torch.zeros(B, T, D, dtype=torch.bool)
            y = {"gt": gtw, "outpainting_mask": maskw, "clip_idx": ii}       # beat:1003-1006
            if ii > 0:
                maskw[:, :OV] = True
                gtw[:, :OV] = prev[:, -OV:]
                y["previous_noisy_tail"] = tail                              # beat:1022-1023
            out = diff.ddim_sample_loop(model, (B, T, D), clip_denoised=False, progress=False, model_kwargs=dict(y=y, pe_type="pe_sinu"))
            prev, tail = out["sample"], out["saved_noisy_tail"]              # beat:1026-1028
            samples.append(prev.numpy().copy())
    np.savez(os.path.join(HERE, "sampler_same_overlap_noisy_toy.npz"), seed=23, samples=np.stack(samples), calls=len(model.calls))
    print("same_overlap_noisy: calls", len(model.calls), "absmax", float(np.abs(np.stack(samples)).max()))


if __name__ == "__main__":
    main()
