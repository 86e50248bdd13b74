"""Multi-rank parity: clip-sharded sampling + ONE all-gather must reproduce the single-GPU result bit for bit (samples are
independent).  With >= 2 GPUs: one rank per GPU over NCCL (the product configuration).  On a single-GPU lease the two
ranks share cuda:0 and exchange through gloo -- NCCL refuses two ranks on one device -- which still exercises sharding, per-rank
engines and the gather on real kernels."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, q):
    import torch.distributed as dist
    from diffsheg_b200 import FusedSpacedDiffusion, FusedUniDiffuser, generate_batch, get_named_beta_schedule, space_timesteps, synth
    from diffsheg_b200.dist import gather_motion, shard_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    multi = torch.cuda.device_count() >= world
    dev_id = rank if multi else 0
    torch.cuda.set_device(dev_id)
    if multi:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev_id))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.make_cfg("show")
        sd = synth.make_state_dict(cfg, seed=1)
        B, T, Dm = 6, 88, cfg["net_dim_pose"]
        inp = synth.make_inputs(cfg, B, T, seed=2)
        opt = synth.make_opt(cfg)

        def sample(mel, hub, pid, x_T, dev):
            eng = FusedUniDiffuser(sd, cfg, precision="bf16", max_batch=mel.shape[0], max_frames=T, device=dev)
            diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000))
            return generate_batch(opt, eng, diff, mel.cuda(dev), pid.cuda(dev), Dm, {"pretrain_aud_feat": hub.cuda(dev)}, {}, noise=x_T)

        mel, hub, pid, x_T = shard_batch([inp["mel"], inp["hubert"], inp["person_id"], inp["x_T"]], rank, world)
        out = gather_motion(sample(mel, hub, pid, x_T, dev_id), B)    # the single collective of the path
        ok = True
        if rank == 0:
            full = sample(inp["mel"], inp["hubert"], inp["person_id"], inp["x_T"], 0)
            ok = bool(torch.equal(out, full))
        q.put((rank, ok, tuple(out.shape)))
    finally:
        dist.destroy_process_group()


def test_sharded_sampling_equals_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == (6, 88, 232)
