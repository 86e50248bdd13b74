"""Parity report (run on the B200 box): how far is every precision mode of the CUDA path from the reference arithmetic?

For each case the fp64 evaluation of the oracle (the reference's op stream in double precision) is the truth; printed next to
each other are

  |mode - fp64|          for every engine precision mode, and
  |oracle_fp32 - fp64|   the reference's OWN fp32 rounding noise on the same inputs (eager torch fp32, TF32 off)

each as relmax (global), per_channel (every pose channel on its own scale) and rel_rms (tests/parity_util.py).  SURVEY 8(d)
asks for the looser-than-fp32 modes to be justified as  |new - fp64| <~ k * |ref_fp32 - fp64| : k is printed per case.
Cases: one denoiser call (SHOW/CFG, BEAT), the 25-step DDIM loop, and the HEADLINE size (config 2: SHOW B=950 ddim25 CFG 1.25)
against the fp32 oracle run eagerly on the same GPU with the same injected x_T.

  python scripts/parity_report.py [--full-batch 950] [--modes fp32,bf16] > gpurun_out/parity_report.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from diffsheg_b200 import FusedSpacedDiffusion, FusedUniDiffuser, get_named_beta_schedule, space_timesteps, synth  # noqa: E402
from oracle import diffusion as odiff  # noqa: E402
from oracle.denoiser import unidiffuser_forward  # noqa: E402
from parity_util import parity_metrics  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def loop_ours(eng, cfg, inp, prec, x_T):
    opt = synth.make_opt(cfg, ddim=True)
    diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000), precision=prec)
    B, T = x_T.shape[:2]
    kw = dict(audio_emb=inp["mel"], length=None, person_id=inp["person_id"], add_cond={"pretrain_aud_feat": inp["hubert"]},
              y={}, pe_type="pe_sinu")
    return diff.ddim_sample_loop(eng, (B, T, cfg["net_dim_pose"]), noise=x_T, clip_denoised=False, model_kwargs=kw)


def loop_oracle(sd_c, cfg, inp, x_T, dtype):
    d = odiff.OracleDiffusion(1000, "ddim25")
    den = odiff.make_denoise(sd_c, cfg, inp["mel"], inp["person_id"], inp["hubert"], dtype=dtype)
    with torch.no_grad():
        return d.ddim_sample_loop(den, tuple(x_T.shape), y={}, noise=x_T, device="cuda", dtype=dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="fp32,bf16")
    ap.add_argument("--full-batch", type=int, default=950)
    ap.add_argument("--fp64-batch", type=int, default=16, help="batch of the fp64-justified loop / call cases")
    args = ap.parse_args()
    modes = args.modes.split(",")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rep = {"modes": modes, "cases": {}}
    for name, T, t_orig, a, b in (("show", 88, 480, 1.8, 1.5), ("beat", 34, 960, 6.2, 6.1), ("show", 88, 0, 1.0001, 0.0101)):
        B = args.fp64_batch
        cfg = synth.make_cfg(name)
        sd = synth.make_state_dict(cfg, seed=1)
        sd_c = {k: v.cuda() for k, v in sd.items()}
        inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, B, T, seed=2).items()}
        ts = torch.full((B,), t_orig, dtype=torch.long, device="cuda")
        ab = (torch.tensor(a, device="cuda"), torch.tensor(b, device="cuda"))
        with torch.no_grad():
            w64 = unidiffuser_forward(sd_c, cfg, inp["x_T"], ts, ab, inp["mel"], inp["person_id"], inp["hubert"], dtype=torch.float64)
            w32 = unidiffuser_forward(sd_c, cfg, inp["x_T"], ts, ab, inp["mel"], inp["person_id"], inp["hubert"], dtype=torch.float32)
        case = {"oracle_fp32_vs_fp64": parity_metrics(w32, w64)}
        for prec in modes:
            eng = FusedUniDiffuser(sd, cfg, precision=prec, max_batch=B, max_frames=T)
            eng.prepare_window(inp["mel"], inp["hubert"], inp["person_id"])
            got = eng.denoise(inp["x_T"], t_orig, a, b)
            torch.cuda.synchronize()
            case[prec + "_vs_fp64"] = parity_metrics(got, w64)
            case[prec + "_vs_oracle_fp32"] = parity_metrics(got, w32)
            case[prec + "_k"] = {k: case[prec + "_vs_fp64"][k] / max(case["oracle_fp32_vs_fp64"][k], 1e-30) for k in case[prec + "_vs_fp64"]}
            del eng
        rep["cases"][f"call_{name}_B{B}_T{T}_t{t_orig}"] = case
        log("call", name, t_orig, json.dumps(case))
        if t_orig == 0:
            continue
        # ---- 25-step DDIM loop, fp64 truth
        w64 = loop_oracle(sd_c, cfg, inp, inp["x_T"], torch.float64)
        w32 = loop_oracle(sd_c, cfg, inp, inp["x_T"], torch.float32)
        case = {"oracle_fp32_vs_fp64": parity_metrics(w32, w64)}
        for prec in modes:
            eng = FusedUniDiffuser(sd, cfg, precision=prec, max_batch=B, max_frames=T)
            got = loop_ours(eng, cfg, inp, prec, inp["x_T"])
            torch.cuda.synchronize()
            case[prec + "_vs_fp64"] = parity_metrics(got, w64)
            case[prec + "_vs_oracle_fp32"] = parity_metrics(got, w32)
            case[prec + "_k"] = {k: case[prec + "_vs_fp64"][k] / max(case["oracle_fp32_vs_fp64"][k], 1e-30) for k in case[prec + "_vs_fp64"]}
            del eng
        rep["cases"][f"ddim25_loop_{name}_B{B}_T{T}"] = case
        log("loop", name, json.dumps(case))
    # ---- the headline size: config 2 against the fp32 oracle run eagerly on this GPU (same x_T)
    if args.full_batch:
        B, T = args.full_batch, 88
        cfg = synth.make_cfg("show")
        sd = synth.make_state_dict(cfg, seed=1)
        sd_c = {k: v.cuda() for k, v in sd.items()}
        inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, B, T, seed=2).items()}
        t0 = time.time()
        w32 = loop_oracle(sd_c, cfg, inp, inp["x_T"], torch.float32)
        torch.cuda.synchronize()
        case = {"oracle_seconds": time.time() - t0}
        for prec in modes:
            if prec == "fp32" and B > 256:
                continue   # the SIMT parity engine is not a performance path: checked at the small sizes above
            eng = FusedUniDiffuser(sd, cfg, precision=prec, max_batch=B, max_frames=T)
            got = loop_ours(eng, cfg, inp, prec, inp["x_T"])
            torch.cuda.synchronize()
            case[prec + "_vs_oracle_fp32"] = parity_metrics(got, w32)
            del eng
        rep["cases"][f"ddim25_loop_show_B{B}_T{T}_full_size"] = case
        log("full", json.dumps(case))
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
