"""Whole-engine emulation in the CTA-pair regime of the tcgen05 engine (rows >= 4096: cta_group::2 GEMMs, the full-row LayerNorm /
modulate / SiLU epilogue of ffn.linear2 on pairs): SHOW B = 50, T = 88 under CFG = 8800 rows (the size of tests/test_variants_gpu.py::test_variant_denoise_at_a_cta_pair_row_count_matches_oracle), one layer per net, bf16, against the
fp32 oracle.  Offline evidence (minutes per case) for the shipped configuration and for the cond_projection variants, whose pair-regime
GEMM instantiations (e.g. LayerNorm fold + SiLU on K = 512 pairs) no shipped path launches.

    python scripts/emu_pair_regime.py > profiles/r02/emu/pair_regime_emulated_engine.txt
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu  # noqa: E402
from diffsheg_b200 import synth  # noqa: E402
from oracle.denoiser import unidiffuser_forward  # noqa: E402
from parity_util import fmt, parity_metrics  # noqa: E402

if __name__ == "__main__":
    B, T = 50, 88      # 4400 conditional rows: the feat_proj GEMMs (M = rows of the conditional half) take the pair forms too
    for cp, cr in (("mlp_includeX", True), ("mlp_excludeX", True), ("linear_includeX", True), ("linear_excludeX", True), ("mlp_includeX", False)):
        cfg = synth.make_cfg("show", num_layers=1, cond_projection=cp, cond_residual=cr)
        sd = synth.make_state_dict(cfg, seed=1)
        inp = synth.make_inputs(cfg, B, T, seed=5)
        t0 = time.time()
        eng = emu.EmuEngine(sd, cfg, precision="bf16", max_batch=B, max_frames=T, sms=8)
        eng.prepare_window(inp["mel"], inp["hubert"], inp["person_id"])
        got = eng.denoise(inp["x_T"], 480, 1.8, 1.5)
        n = eng.launch_count()
        eng.close()
        ts = torch.full((B,), 480, dtype=torch.long)
        with torch.no_grad():
            want = unidiffuser_forward(sd, cfg, inp["x_T"], ts, (torch.tensor(1.8), torch.tensor(1.5)), inp["mel"], inp["person_id"], inp["hubert"])
        print(f"show B{B} T{T} (2 x {B * T} rows) 1 layer bf16 {cp} cond_residual={cr}: {fmt(parity_metrics(got, want))}  ({time.time() - t0:.0f} s, {n} launches)", flush=True)
