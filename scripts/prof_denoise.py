"""Profiling driver: a few denoiser calls at a given batch (run under ncu on the GPU box)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsheg_b200 import FusedUniDiffuser, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=950)
ap.add_argument("--calls", type=int, default=3)
ap.add_argument("--name", default="show")
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
cfg = synth.make_cfg(a.name)
sd = synth.make_state_dict(cfg, seed=1)
T = cfg["n_poses"]
eng = FusedUniDiffuser(sd, cfg, precision=a.precision, max_batch=a.batch, max_frames=T)
inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, a.batch, T, seed=2).items()}
eng.prepare_window(inp["mel"], inp["hubert"], inp["person_id"])
out = torch.empty_like(inp["x_T"])
for i in range(a.calls):
    eng.denoise(inp["x_T"], 480, 1.8, 1.5, out=out)
torch.cuda.synchronize()
print("launches", eng.launch_count(), float(out.abs().mean()))
