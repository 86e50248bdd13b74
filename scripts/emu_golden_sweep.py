"""Whole-engine emulation at the FULL golden sizes (8 layers, SHOW B=2 T=88 under CFG / BEAT B=2 T=34): the emulated engine
(tests/emu/emu_engine.cpp: engine.cu + every kernel on the CPU) against the outputs of the REAL reference in tests/golden/ --
the same comparison tests/test_variants_gpu.py::test_variant_denoise_matches_reference_golden and
tests/test_gpu_parity.py::test_denoise_matches_reference_golden make on the B200.  Offline evidence (about a minute per case);
the -m "not gpu" suite runs the small versions (tests/test_emu_engine.py).

    python scripts/emu_golden_sweep.py [shipped|variants] > profiles/r02/emu/<name>.txt
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu  # noqa: E402
from diffsheg_b200 import synth  # noqa: E402
from parity_util import fmt, parity_metrics  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(name, prec, B, T, t_orig, a, b, want, **over):
    cfg = synth.make_cfg(name, **over)
    sd = synth.make_state_dict(cfg, seed=1)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    t0 = time.time()
    eng = emu.EmuEngine(sd, cfg, precision=prec, max_batch=B, max_frames=T)
    eng.prepare_window(inp["mel"], inp["hubert"], inp["person_id"])
    out = eng.denoise(inp["x_T"], int(t_orig), float(a), float(b))
    n = eng.launch_count()
    eng.close()
    return f"{fmt(parity_metrics(out, torch.from_numpy(want)))}  ({time.time() - t0:.0f} s, {n} launches)"


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "variants"
    if what == "shipped":
        for name, B, T, t_resp in (("show", 2, 88, 12), ("show", 3, 84, 0), ("beat", 2, 34, 24), ("beat", 1, 30, 3)):
            g = np.load(os.path.join(GOLDEN, f"denoise_{name}_B{B}_T{T}_t{t_resp}.npz"))
            for prec in ("fp32", "tf32", "bf16"):
                print(f"shipped {name} B{B} T{T} t{t_resp} {prec}: " + run(name, prec, B, T, g["t_orig"], g["a"], g["b"], g["eps"]), flush=True)
    else:
        g = np.load(os.path.join(GOLDEN, "denoise_variants.npz"))
        for name in ("beat", "show"):
            B, T, _, t_orig, a, b = g[name + "_consts"]
            for cp in ("mlp_includeX", "linear_includeX", "mlp_excludeX", "linear_excludeX"):
                for cr in (True, False):
                    if cp == "mlp_includeX" and cr:
                        continue
                    for prec in ("fp32", "tf32", "bf16"):
                        want = g[f"{name}_{cp}_{'res' if cr else 'nores'}"]
                        print(f"{name} {cp} cond_residual={cr} {prec}: " +
                              run(name, prec, int(B), int(T), t_orig, a, b, want, cond_projection=cp, cond_residual=cr), flush=True)
