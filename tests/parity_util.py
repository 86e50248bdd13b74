"""Error metrics shared by the GPU parity tests, smoke() and scripts/parity_report.py (test infrastructure).

``relmax`` alone (max|d| / max|want| over the WHOLE tensor) lets a channel whose magnitude is 1 % of the largest one be
100 % wrong; the gates therefore also bound

* ``per_channel`` : max over the pose channels c of  max|d[..., c]| / max|want[..., c]|   (every channel on its own scale)
* ``rel_rms``     : ||d||_2 / ||want||_2                                                    (typical, not worst-case, error)
"""
import torch


def parity_metrics(got, want):
    g, w = got.detach().double().cpu(), want.detach().double().cpu()
    d = (g - w).abs()
    ch_err = d.reshape(-1, d.shape[-1]).amax(0)
    ch_ref = w.abs().reshape(-1, w.shape[-1]).amax(0)
    return {"relmax": float(d.max() / (w.abs().max() + 1e-30)),
            "per_channel": float((ch_err / (ch_ref + 1e-30)).max()),
            "rel_rms": float(d.pow(2).sum().sqrt() / (w.pow(2).sum().sqrt() + 1e-30))}


def fmt(m):
    return " ".join(f"{k}={v:.3e}" for k, v in m.items())


def check(m, gate, what=""):
    """gate: dict with the same keys; every metric must stay below its bound."""
    bad = {k: (m[k], gate[k]) for k in gate if not m[k] < gate[k]}
    assert not bad, f"{what}: parity gate exceeded {bad} (all metrics: {fmt(m)})"
