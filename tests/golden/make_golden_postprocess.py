"""Golden vectors for the output post-processing, produced by the REFERENCE's own functions.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_postprocess.py
* BEAT axis-angle branch, trainers/ddpm_beat_trainer.py:1056-1062, with datasets/rotation_converter.py imported by file
  path (the ``datasets`` package itself needs lmdb, which is absent here);
* SHOW ``inv_standardize`` (datasets/show.py:157-162; the class needs lmdb) evaluated as the one numpy expression it is.
"""
import importlib.util
import os

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))


def load_rc():
    spec = importlib.util.spec_from_file_location("ref_rotation_converter", "/root/reference/datasets/rotation_converter.py")
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    return rc


def main():
    rc = load_rc()
    g = torch.Generator().manual_seed(11)
    B, T, C = 3, 34, 141                                   # BEAT gesture: 47 joints (runner.py:127-145)
    x = torch.randn(B, T, C, generator=g)
    std_aa = torch.rand(C, generator=g) * 0.3 + 0.05
    mean_aa = torch.randn(C, generator=g) * 0.2
    mean_pose = torch.randn(C, generator=g) * 20
    std_pose = torch.rand(C, generator=g) * 15 + 1
    x[0, 0, :3] = -mean_aa[:3] / std_aa[:3]                # denormalises to (almost) the zero rotation: small-angle branch
    # --- verbatim call sequence of beat:1056-1062 on torch CPU tensors
    denorm_out = x * std_aa + mean_aa
    euler_out = rc.axis_angle_to_euler_angles(denorm_out.reshape(B, T, C // 3, 3)).reshape(B, T, C)
    euler_out = euler_out * (180 / np.pi)
    out_motions = (euler_out - mean_pose) / std_pose
    np.savez_compressed(os.path.join(OUT, "postprocess_beat.npz"), x=x.numpy(), std_aa=std_aa.numpy(), mean_aa=mean_aa.numpy(),
                        mean_pose=mean_pose.numpy(), std_pose=std_pose.numpy(), euler_deg=euler_out.numpy(),
                        out_motions=out_motions.numpy())
    # --- SHOW: show.py:159 on a [B,T,232] sample, then the split of show:920-921
    D = 232
    m = torch.randn(2, 88, D, generator=g).numpy()
    mean = (torch.randn(D, generator=g) * 0.5).numpy()
    std = (torch.rand(D, generator=g) + 0.1).numpy()
    inv = m * std + mean
    np.savez_compressed(os.path.join(OUT, "postprocess_show.npz"), x=m, mean=mean, std=std, inv=inv, split_pos=129)
    print("postprocess goldens written")


if __name__ == "__main__":
    main()
