"""Device-side output post-processing: what the trainers do to the sampled motion AFTER the sample loop.

The reference copies every window to the host, concatenates in numpy and then de-standardises (SHOW,
trainers/ddpm_show_trainer.py:896-921) or de-normalises + converts axis-angle to Euler degrees on torch-CPU (BEAT
``--axis_angle``, trainers/ddpm_beat_trainer.py:1056-1062).  Here the stitched sample never leaves the GPU
(``generate_long``), the arithmetic runs in two fused kernels (csrc/postprocess.cuh) and each result array is copied to
pinned host memory once.  File writers (``np.save``, BVH, ARKit json: show:923-931, beat:1064-1086) stay with the caller.
"""
import torch

from . import _lib
from .engine import _ptr, _stream


def _stat(v, n, dev, what):
    t = torch.as_tensor(v, dtype=torch.float32).reshape(-1).to(dev).contiguous()
    if t.numel() != n:
        raise ValueError(f"{what}: expected {n} values, got {t.numel()}")
    return t


def _rows(motion):
    """2-D [rows, channels] view (unit channel stride, uniform row stride); copies only when the layout forces it."""
    if motion.dim() < 2 or motion.dtype != torch.float32 or not motion.is_cuda:
        raise ValueError("motion must be a CUDA fp32 tensor [..., channels]")
    flat = motion.reshape(-1, motion.shape[-1])
    return flat if flat.stride(1) == 1 else flat.contiguous()


def inv_standardize(motion, mean, std, columns=None):
    """``dataset.inv_standardize(out_motions, mean, std)`` (datasets/show.py:157-162) on the device.

    ``columns=(lo, hi)`` restricts the operation to that channel window of ``motion`` (the gesture / expression split of
    show:920-921) with ``mean`` / ``std`` of length ``hi - lo``; the result is a new dense ``[..., hi - lo]`` tensor."""
    flat = _rows(motion)
    lo, hi = columns if columns is not None else (0, flat.shape[1])
    if not (0 <= lo < hi <= flat.shape[1]):
        raise ValueError(f"bad column window {columns} for {flat.shape[1]} channels")
    D = hi - lo
    dev = flat.device
    out = torch.empty(flat.shape[0], D, device=dev, dtype=torch.float32)
    m, s = _stat(mean, D, dev, "mean"), _stat(std, D, dev, "std")
    with torch.cuda.device(dev):
        src = flat[:, lo:]
        _lib.check(_lib.lib().dsheg_inv_standardize(_ptr(src), flat.stride(0), _ptr(m), _ptr(s), _ptr(out), D, flat.shape[0], D,
                                                    _stream(dev)), None, "inv_standardize")
    return out.reshape(*motion.shape[:-1], D)


def axis_angle_to_euler(motion, mean_aa, std_aa, mean_pose, std_pose, channels=None):
    """beat:1056-1062 on the device.  ``motion [..., >= C]`` holds the normalised axis-angle gesture in its first
    ``channels`` (= ``opt.split_pos`` = 3 * joints) columns.  Returns ``(euler_deg, out_motions)``, both ``[..., C]``:
    the de-normalised XYZ Euler angles in degrees (the BVH input, beat:1076) and their re-normalised form (beat:1061)."""
    flat = _rows(motion)
    C = channels if channels is not None else flat.shape[1]
    if C % 3 or not (0 < C <= flat.shape[1]):
        raise ValueError(f"axis-angle channel count must be a positive multiple of 3 and <= {flat.shape[1]}, got {C}")
    dev = flat.device
    st = [_stat(v, C, dev, n) for v, n in ((mean_aa, "mean_aa"), (std_aa, "std_aa"), (mean_pose, "mean_pose"), (std_pose, "std_pose"))]
    euler = torch.empty(flat.shape[0], C, device=dev, dtype=torch.float32)
    out = torch.empty_like(euler)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dsheg_beat_axis_angle_to_euler(_ptr(flat), flat.stride(0), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]),
                                                             _ptr(st[3]), _ptr(euler), _ptr(out), flat.shape[0], C, _stream(dev)),
                   None, "axis_angle_to_euler")
    shape = (*motion.shape[:-1], C)
    return euler.reshape(shape), out.reshape(shape)


def resample_features(feat, n_out):
    """show:1082 / datasets/show.py:98: ``F.interpolate(feat.swapaxes(-1,-2), size=n_out, mode='linear', align_corners=True)
    .swapaxes(-1,-2)`` for HuBERT features ``[B, n_in, C]`` (or ``[n_in, C]``) on the device, without the two transposes."""
    if not (torch.is_tensor(feat) and feat.is_cuda):
        raise ValueError("resample_features expects a CUDA tensor")
    squeeze = feat.dim() == 2
    x = (feat.unsqueeze(0) if squeeze else feat).to(torch.float32).contiguous()
    B, n_in, C = x.shape
    if C % 4:
        raise ValueError(f"channel count must be a multiple of 4, got {C}")
    out = torch.empty(B, int(n_out), C, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dsheg_resample_linear(_ptr(x), _ptr(out), B, n_in, int(n_out), C, _stream(x.device)), None, "resample_linear")
    return out[0] if squeeze else out


def _to_host(t):
    """One D2H into pinned memory; returns a numpy view (the arrays the trainers hand to np.save)."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def finish_show(opt, motion, motion_mean, motion_std):
    """show:909-921 for the unidiffuser configuration: returns ``(out_motions, out_expression)`` as numpy arrays,
    de-standardised and split at ``opt.split_pos``."""
    sp = opt.split_pos
    mean = torch.as_tensor(motion_mean, dtype=torch.float32).reshape(-1)
    std = torch.as_tensor(motion_std, dtype=torch.float32).reshape(-1)
    ges = inv_standardize(motion, mean[:sp], std[:sp], columns=(0, sp))
    exp = inv_standardize(motion, mean[sp:], std[sp:], columns=(sp, motion.shape[-1]))
    return _to_host(ges), _to_host(exp)


def finish_beat(opt, motion, mean_aa=None, std_aa=None, mean_pose=None, std_pose=None):
    """beat:1050-1062: split at ``opt.split_pos``; with ``opt.axis_angle`` also the Euler conversion.  Returns a dict with
    ``motions`` (what beat:1068/1078 saves), ``expression`` and, for axis-angle models, ``axis_angle`` (beat:1053-1055
    saves it) and ``euler_deg`` (``out_denorm_euler`` of beat:1075)."""
    sp = opt.split_pos
    res = {"expression": _to_host(motion[..., sp:].contiguous())}
    if getattr(opt, "axis_angle", False):
        euler, out = axis_angle_to_euler(motion, mean_aa, std_aa, mean_pose, std_pose, channels=sp)
        res.update(axis_angle=_to_host(motion[..., :sp].contiguous()), euler_deg=_to_host(euler), motions=_to_host(out))
    else:
        res["motions"] = _to_host(motion[..., :sp].contiguous())
    return res
