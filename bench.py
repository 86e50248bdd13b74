#!/usr/bin/env python
"""Benchmark of the DiffSHEG sampling hot path (BASELINE.json metric: motion-frames/sec, ddim25,
n_poses=88, bs=950).  One "step" = one full ddim_sample_loop over one batch of synthetic input
(25 denoiser calls + 25 fused DDIM updates).  Contract: see the task statement / DESIGN.md section 6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "motion-frames/sec (ddim25, n_poses=88, bs=950)"
UNIT = "frames/s"
CANON_FLOP_PER_FRAME_CALL = 258.83e6  # SHOW + CFG, FlopCounterMode on the reference op stream
# algorithmic bytes (bf16 A + output + residual, W negligible) of one layer's GEMMs at R = 167 200 rows, cond half 83 600:
# feat1 (2 KB in + 2 KB out per cond row), feat2 (2+1+1), qkv (1+3), sa_out (1+1+1), ffn1 (1+2), ffn2 (2+1), ffn_out (1+1+1)
ALGO_GEMM_BYTES_PER_LAYER = 83600 * (4096 + 4096) + 167200 * (4096 + 3072 + 3072 + 3072 + 3072)


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            m = json.load(open(f))
            p.update(hbm_gbs=float(m["hbm_gbs"]), bf16_tflops=float(m["bf16_tflops"]),
                     bf16_tflops_sustained=float(m.get("bf16_tflops_sustained", m["bf16_tflops"])),
                     source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def cpu_threads():
    """Threads for the torch-CPU baseline: every core up to 32.  On the 128-core GPU host, torch's intra-op pools get
    SLOWER beyond that on these matrix sizes (measured: 128 threads -> 3.2 frames/s at B=4; see profiles/r01)."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("DSHEG_CPU_THREADS", "32"))))


def oracle_cpu_frames_per_s(cfg, B, steps, warmup, threads):
    """The reference's CPU implementation of the path (oracle port: same torch fp32 op stream), timed on host cores."""
    from diffsheg_b200 import synth
    from oracle import diffusion as odiff
    torch.set_num_threads(threads)
    sd = synth.make_state_dict(cfg, seed=1)
    T = cfg["n_poses"]
    inp = synth.make_inputs(cfg, B, T, seed=2)
    d = odiff.OracleDiffusion(1000, "ddim25")
    den = odiff.make_denoise(sd, cfg, inp["mel"], inp["person_id"], inp["hubert"])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            d.ddim_sample_loop(den, (B, T, cfg["net_dim_pose"]), y={}, noise=inp["x_T"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return B * T / (ms / 1e3), ms


def run_reference(args):
    """--impl reference: the reference is pure Python/torch and cannot travel to the GPU box, so this arm times the
    oracle port (identical torch-CPU op stream, pinned against the real reference's outputs) on all host cores."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from diffsheg_b200 import synth
    cfg = synth.make_cfg("show")
    cores = cpu_threads()
    B = args.ref_batch
    fps, ms = oracle_cpu_frames_per_s(cfg, B, args.steps, args.warmup, cores)
    sample = f"B={B} of the bs=950 workload per step (SHOW T=88 CFG 1.25 ddim25, all 25 calls); frames/s is batch-linear on CPU"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SHOW n_poses=88 ddim25 cond_scale=1.25 batch=950 (configs[1])", "sample": sample},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ref_cuda(args):
    """R-cuda row of BASELINE.md: the reference's eager torch op stream (oracle port) on the SAME GPU, fp32, torch default
    matmul precision (TF32 off) and with TF32 on -- the denominator of north_star's >= 10x target.  Informational."""
    from diffsheg_b200 import synth
    from oracle import diffusion as odiff
    cfg = synth.make_cfg("show")
    B, T = args.ref_cuda, cfg["n_poses"]
    sd = {k: v.cuda() for k, v in synth.make_state_dict(cfg, seed=1).items()}
    inp = {k: v.cuda() for k, v in synth.make_inputs(cfg, B, T, seed=2).items()}
    d = odiff.OracleDiffusion(1000, "ddim25")
    den = odiff.make_denoise(sd, cfg, inp["mel"], inp["person_id"], inp["hubert"])
    rows = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            d.ddim_sample_loop(den, (B, T, cfg["net_dim_pose"]), y={}, noise=inp["x_T"], device="cuda")  # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            d.ddim_sample_loop(den, (B, T, cfg["net_dim_pose"]), y={}, noise=inp["x_T"], device="cuda")
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows[name] = {"ms_per_step": ms, "frames_per_s": B * T / (ms / 1e3)}
    print(json.dumps({"impl": "reference-op-stream torch-cuda eager (oracle port)", "metric": METRIC, "unit": UNIT, "batch": B,
                      "note": "no generate_src_mask host syncs (SURVEY F8): an upper bound on the real reference", "rows": rows}))


def run_ours(args):
    from diffsheg_b200 import FusedSpacedDiffusion, FusedUniDiffuser, generate_batch, get_named_beta_schedule, space_timesteps, synth
    from diffsheg_b200.dist import gather_motion
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = synth.make_cfg("show")
    B, T, Dm = args.batch, cfg["n_poses"], cfg["net_dim_pose"]
    sd = synth.make_state_dict(cfg, seed=1)
    eng = FusedUniDiffuser(sd, cfg, precision=args.precision, max_batch=B, max_frames=T, device=local)
    opt = synth.make_opt(cfg)
    diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000),
                                precision=args.precision)
    inp = synth.make_inputs(cfg, B, T, seed=100 + rank)
    host = {k: inp[k].pin_memory() for k in ("mel", "hubert", "person_id")}
    devin = {k: v.to(dev) for k, v in host.items()}
    out_host = torch.empty(B, T, Dm).pin_memory()
    total_B = B * world

    def step_resident():
        out = generate_batch(opt, eng, diff, devin["mel"], devin["person_id"], Dm, {"pretrain_aud_feat": devin["hubert"]}, {})
        return gather_motion(out, total_B) if world > 1 else out

    def step_e2e():
        mel = host["mel"].to(dev, non_blocking=True)
        hub = host["hubert"].to(dev, non_blocking=True)
        pid = host["person_id"].to(dev, non_blocking=True)
        out = generate_batch(opt, eng, diff, mel, pid, Dm, {"pretrain_aud_feat": hub}, {})
        out_host.copy_(out, non_blocking=True)   # every rank reads its own result back; rank 0 also gathers
        return gather_motion(out, total_B) if world > 1 else out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        barrier()
        return ms

    for _ in range(args.warmup):
        step_resident()
    launches0, steps0 = eng.launch_count(), diff.step_launches
    with ClockSampler(local) as cs:
        ms = timed(step_resident, args.steps)
    launches = (eng.launch_count() - launches0) + (diff.step_launches - steps0)
    clocks = cs.summary()
    frames = total_B * T
    value = frames / (ms / 1e3)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = frames / (ms_e2e / 1e3)
    h2d = sum(v.numel() * 4 for v in host.values())
    d2h = out_host.numel() * 4

    # ---- roofline pass: CUDA events around every GEMM / attention / row-wise launch of one step (after the timed region)
    pk = peaks()
    eng.profile_begin()
    step_resident()
    prof = eng.profile_end()
    g, at = prof["gemm"], prof["attention"]
    gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    attn_gbs = at["work"] / (at["ms"] * 1e-3) / 1e9 if at["ms"] > 0 else 0.0
    roof = {"kernel": "gemm_tc_kernel (tcgen05 bf16, all per-step GEMMs)" if args.precision == "bf16" else "gemm_simt_kernel (fp32)",
            "bound": "tensor", "achieved": gemm_tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": gemm_tflops / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": pk["source"] + ", sustained",
            "launches_per_step": g["count"], "ms_per_step": g["ms"], "share_of_step": g["ms"] / ms if world == 1 else None,
            "executed_flops_per_step": g["work"]}
    roof_attn = {"kernel": "attn_%s kernel (linear attention + LN/modulate/SiLU)" % (os.environ.get("DSHEG_ATTN") or "v3"), "bound": "hbm", "achieved": attn_gbs,
                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": attn_gbs / pk["hbm_gbs"], "traffic": None,
                 "peak_source": pk["source"], "launches_per_step": at["count"], "ms_per_step": at["ms"],
                 "share_of_step": at["ms"] / ms if world == 1 else None}
    if rank != 0:
        return
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (one layer's 7 GEMM launches)
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r01", "final_gemm_ncu_summary.json")))
        roof["traffic"] = 1e6 * sum(k["dram_read_MB"] + k["dram_write_MB"] for k in ncu) / len(ncu)
        roof["traffic_note"] = ("mean dram__bytes_read+write per launch over the 7 GEMMs of one layer (ncu --set full, "
                                "profiles/r01/final_gemm_ncu_summary.json); algorithmic operand+output bytes of the same launches: "
                                f"{ALGO_GEMM_BYTES_PER_LAYER / 7 / 1e6:.0f} MB per launch")
    except Exception:
        pass
    try:   # same for the attention kernel (default kernel attn_v3; one launch, SHOW B=950 CFG)
        na = json.load(open(os.path.join(ROOT, "profiles", "r01", "final_attn_ncu_summary.json")))
        if not os.environ.get("DSHEG_ATTN"):
            roof_attn["traffic"] = 1e6 * (na["dram__bytes_read.sum"][0] + na["dram__bytes_write.sum"][0])
            roof_attn["traffic_note"] = ("dram__bytes_read+write of one launch (ncu --set full, profiles/r01/final_attn_ncu_summary.json); "
                                         f"algorithmic bytes of the same launch: {at['work'] / max(at['count'], 1) / 1e6:.0f} MB")
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "SHOW n_poses=88 ddim25 cond_scale=1.25 batch=950 per GPU (configs[1]); one step = one "
                                   "ddim_sample_loop = 25 denoiser calls (CFG pair) + 25 fused DDIM updates",
                       "per_gpu_batch": B, "global_batch": total_B, "frames_per_step": frames, "parallelism": f"dp{world}",
                       "l2": "per-call activations (>1 GB) and conditioning (385 MB) exceed the 126 MB L2; no explicit flush",
                       "precision": args.precision, "final_all_gather": world > 1,
                       # experiment switches in effect (empty = the shipped defaults), so variant runs describe themselves
                       "switches": {k: v for k, v in os.environ.items() if k.startswith("DSHEG_") and v}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world,
                    "api": "diffsheg_b200.generate_batch (trainers' generate_batch seam) with pinned host mel/HuBERT/person-id in, pinned host motion out"},
            "gpu_launches": launches,
            "model_tflops": frames * 25 * CANON_FLOP_PER_FRAME_CALL / (ms / 1e3) / 1e12,
            "roofline": roof, "roofline_attention": roof_attn,
            "rowwise": {"ms_per_step": prof["rowwise"]["ms"], "gbs": prof["rowwise"]["work"] / max(prof["rowwise"]["ms"], 1e-9) / 1e6}}
    if world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads()
        fps, cms = oracle_cpu_frames_per_s(cfg, args.ref_batch, 1, 1, cores)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"oracle port (reference torch-CPU op stream), B={args.ref_batch} of the 950-batch, "
                                          f"full 25-step loop, 1 warm-up + 1 timed ({cms / 1e3:.1f} s)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=950, help="per-GPU batch (BASELINE configs[1]: 950)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=8, help="bounded CPU sample of the workload (about 10 s of CPU work per loop)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-cuda", type=int, default=0, help="time the reference op stream (oracle port) eagerly on the GPU at this batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.ref_cuda:
        run_ref_cuda(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
