"""Bandwidth of the post-processing kernels at the BASELINE batch sizes; CUDA events, L2 flushed between runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import diffsheg_b200 as dz  # noqa: E402


def timed(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


x = torch.randn(950, 88, 232, device="cuda")
mean, std = torch.randn(232, device="cuda"), torch.rand(232, device="cuda") + 0.1
ms = timed(lambda: dz.inv_standardize(x, mean, std))
print(f"inv_standardize SHOW B=950: {ms * 1e3:.1f} us, {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s (read + write; includes the output allocation)")
xb = torch.randn(2500, 34, 192, device="cuda")
st = [torch.randn(141, device="cuda") * 0.2, torch.rand(141, device="cuda") * 0.3 + 0.05, torch.randn(141, device="cuda") * 20,
      torch.rand(141, device="cuda") * 15 + 1]
ms = timed(lambda: dz.axis_angle_to_euler(xb, *st, channels=141))
n = 2500 * 34 * 141
print(f"axis_angle_to_euler BEAT B=2500: {ms * 1e3:.1f} us, {3 * n * 4 / ms / 1e6:.0f} GB/s (1 read + 2 writes of [rows,141] fp32)")
