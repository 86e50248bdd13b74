"""First hardware run of the kernels written without GPU access (validated so far only on the CPU emulator, tests/emu):
attention v4 / v5<1,2,4> and the post-processing kernels.  Prints one PASS / FAIL line per item plus the attention
bandwidth of every variant inside a real denoiser call, and writes the same to gpurun_out/first_hw_run.json.

Run by tests/test_zz_first_hw_run.py in a SUBPROCESS (a faulting kernel must not poison the CUDA context of the parity
suite) and by scripts/gpu_round2_first.sh.  Exit code 0 = everything passed."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = ["v4", "v5c1", "v5c2", "v5c4"]   # op-level parity for all four; in-loop agreement below
LOOP_ONLY = ["v5c1+expo", "v5c4+expo",   # + DSHEG_EXPO=1: Q and K numerators with static shifts from the epilogue, attn_v5<CL, 2>
             "v6c2+expo", "v6c1+expo",    # attn_v6 as clusters of two 512-thread CTAs / as one 1024-thread CTA without a cluster
             "v6+expo",                   # attn_v6: 4 warps per head, 64 registers, 32 warps per SM (needs the EXPO numerators)
             "default+aud",               # + DSHEG_ATTN_AUD=1: attn_small.cuh for the audio encoder layer (all 8 heads of 16 at once)
             "default+lnms",              # + DSHEG_FUSE_LNMS=1: ffn.linear2 + LayerNorm / modulate / SiLU in one GEMM (ACT_LNMS)
             "v6+expo+lnms+aud",          # everything at once
             "v5c4+qsoft"]                # + DSHEG_QSOFT=1: Q row-softmax only (superseded by EXPO when that wins)
results = {}


def record(name, ok, **info):
    results[name] = dict(ok=bool(ok), **info)
    print(("PASS " if ok else "FAIL ") + name + (" " + json.dumps(info) if info else ""), flush=True)


def pytest_items():
    env = dict(os.environ, DSHEG_RUN_UNVALIDATED="1")
    items = [("postprocess kernels (tests/test_postprocess.py)", ["tests/test_postprocess.py", "-k", "gpu_"])]
    items += [(f"attention op {v} (test_op_attention_bf16_tensor_core)", ["tests/test_gpu_parity.py", "-k", f"op_attention_bf16 and {v}"])
              for v in VARIANTS]   # one process per variant: a faulting kernel poisons only its own CUDA context
    items += [("GEMM op ACT_EXPO (test_op_linear_exponential_epilogue)", ["tests/test_gpu_parity.py", "-k", "exponential_epilogue"]),
              ("GEMM op ACT_LNMS (test_op_linear_layernorm_modulate_silu_epilogue)", ["tests/test_gpu_parity.py", "-k", "layernorm_modulate_silu_epilogue"])]
    items += [(f"attention op {v} + static-shift numerators", ["tests/test_gpu_parity.py", "-k", f"static_shift_numerators and {v}"])
              for v in ("v5c1", "v5c2", "v5c4", "v6", "v6c2", "v6c1")]
    for name, sel in items:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + sel, cwd=ROOT, env=env,
                               capture_output=True, text=True, timeout=120)
            tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
            record(name, r.returncode == 0, seconds=round(time.time() - t0, 1), tail=tail)
        except subprocess.TimeoutExpired:
            record(name, False, error="timeout")


def in_loop(batch, var):
    """Default attention vs ONE variant inside dsheg_denoise (SHOW, CFG): output agreement and attention GB/s."""
    import torch
    from diffsheg_b200 import FusedUniDiffuser, synth
    cfg = synth.make_cfg("show")
    sd = synth.make_state_dict(cfg, seed=1)
    T = cfg["n_poses"]
    inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, batch, T, seed=2).items()}
    base = None
    for v in (None, var):
        name = f"denoise B={batch} attention={v or 'v3 (default)'}"
        os.environ.pop("DSHEG_ATTN", None)
        os.environ.pop("DSHEG_QSOFT", None)
        os.environ.pop("DSHEG_EXPO", None)
        os.environ.pop("DSHEG_FUSE_LNMS", None)
        os.environ.pop("DSHEG_ATTN_AUD", None)
        if v:
            parts = v.split("+")
            if parts[0] != "default":
                os.environ["DSHEG_ATTN"] = parts[0]
            for flag, env in (("qsoft", "DSHEG_QSOFT"), ("expo", "DSHEG_EXPO"), ("lnms", "DSHEG_FUSE_LNMS"), ("aud", "DSHEG_ATTN_AUD")):
                if flag in parts[1:]:
                    os.environ[env] = "1"
        eng = FusedUniDiffuser(sd, cfg, precision="bf16", max_batch=batch, max_frames=T)
        eng.prepare_window(inp["mel"], inp["hubert"], inp["person_id"])
        out = torch.empty_like(inp["x_T"])
        for _ in range(2):
            eng.denoise(inp["x_T"], 480, 1.8, 1.5, out=out)
        eng.profile_begin()
        for _ in range(2):
            eng.denoise(inp["x_T"], 480, 1.8, 1.5, out=out)
        at = eng.profile_end()["attention"]
        torch.cuda.synchronize()
        gbs = at["work"] / (at["ms"] * 1e-3) / 1e9 if at["ms"] > 0 else 0.0
        if base is None:
            base = out.clone()
            record(name, bool(torch.isfinite(out).all()), attention_gbs=round(gbs), attention_ms=round(at["ms"], 3))
        else:
            err = float((out - base).abs().max() / base.abs().max())
            record(name, err < 2e-2 and bool(torch.isfinite(out).all()), relmax_vs_default=err, attention_gbs=round(gbs),
                   attention_ms=round(at["ms"], 3))
        del eng


def child(var):
    """One variant per process: a faulting kernel poisons only its own CUDA context."""
    # B = 3: single-CTA GEMM variants; B = 24 (4 224 rows): the CTA-pair variants, the only ones ACT_LNMS exists for and the
    # ones ACT_EXPO / ACT_QSOFT run as at the headline batch; DSHEG_FIRST_RUN_BATCH=0: no full-size run
    batches = [3] + ([24] if "+" in var else []) + [int(os.environ.get("DSHEG_FIRST_RUN_BATCH", "950"))]
    for batch in batches:
        if batch <= 0:
            continue
        try:
            in_loop(batch, var)
        except Exception as e:  # noqa: BLE001
            record(f"denoise B={batch} attention={var}", False, error=repr(e)[:300])
            break
    print("RESULTS " + json.dumps(results), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        child(sys.argv[2])
        sys.exit(0)
    pytest_items()
    for var in VARIANTS + LOOP_ONLY:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", var], cwd=ROOT, capture_output=True, text=True,
                               timeout=240)
            got = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULTS ")]
            if got:
                for k, v in json.loads(got[-1][8:]).items():
                    if not (k.endswith("(default)") and k in results):
                        results[k] = v
                        print(("PASS " if v["ok"] else "FAIL ") + k + " " + json.dumps({a: b for a, b in v.items() if a != "ok"}), flush=True)
            else:
                record(f"in-loop {var}", False, rc=r.returncode, tail=(r.stdout + r.stderr).strip().splitlines()[-3:])
        except subprocess.TimeoutExpired:
            record(f"in-loop {var}", False, error="timeout")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "first_hw_run.json"), "w") as f:
        json.dump(results, f, indent=1)
    sys.exit(0 if all(r["ok"] for r in results.values()) else 1)
