"""Host-side multi-rank logic on CPU (gloo, world_size 2): batch sharding + the single final all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsheg_b200.dist import gather_motion, shard_batch, shard_range


def test_shard_range_partitions():
    for n in (1, 2, 7, 950, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 3 * 5, dtype=torch.float32).view(total, 3, 5)
        cond = {"pretrain_aud_feat": full * 2}
        mine, mycond = shard_batch([full, cond], rank, world)
        lo, hi = shard_range(total, rank, world)
        assert mine.shape[0] == hi - lo and torch.equal(mycond["pretrain_aud_feat"], full[lo:hi] * 2)
        out = gather_motion(mine + 1.0, total)   # stands in for the per-rank sample loop
        q.put((rank, bool(torch.equal(out, full + 1.0))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])  # even split -> all_gather_into_tensor; ragged -> padded all_gather
def test_gather_motion_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
