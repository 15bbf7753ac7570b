"""Multi-process (world_size 2, gloo, CPU) tests of the batch-sharding host logic
(advchain_b200/augmentor/sharding.py).  The CUDA kernels cannot run here, so the
per-shard arithmetic is done by the CPU oracle; what is under test is that the three scalar
exchanges of `ShardContext` make two shards reproduce the unsharded run:
  clamp bounds, the Q9 loss weighting (shard losses sum to the global loss and give the global
  per-sample gradients), and the whole-batch norm of the 3-D step rule."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import advchain_oracle as orc
from tests.golden.cases import stage_cfgs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, global_batch, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from advchain_b200.augmentor.sharding import ShardContext
        torch.set_num_threads(1)
        ctx = ShardContext(global_batch)
        sl = ctx.local
        torch.manual_seed(0)                      # every rank builds the same global problem
        size = [global_batch, 1, 24, 32]
        data = torch.rand(*size) * 3 - 1
        logits = torch.randn(global_batch, 4, 24, 32)
        ref = torch.randn(global_batch, 4, 24, 32)
        mask = (torch.rand(global_batch, 1, 24, 32) > 0.2).float().expand(-1, 4, -1, -1)
        types, weights = ["mse", "contour"], [1.0, 0.5]
        res = {"rank": rank, "slice": (sl.start, sl.stop)}
        # 1. clamp bounds
        res["minmax"] = ctx.global_minmax(data[sl])
        res["minmax_ref"] = (float(data.min()), float(data.max()))
        # 2. loss: local value with scaled weights, summed; gradients vs the unsharded gradient
        full = logits.clone().requires_grad_(True)
        l_full = orc.consistency_loss(full, ref, tuple(types), tuple(weights), mask)
        l_full.backward()
        loc = logits[sl].clone().requires_grad_(True)
        w = ctx.scaled_weights(types, weights)
        l_loc = orc.consistency_loss(loc, ref[sl], tuple(types), tuple(w), mask[sl])
        l_loc.backward()
        res["loss_sum"] = float(ctx.global_scalar(l_loc))
        res["loss_ref"] = float(l_full)
        res["grad_err"] = float((loc.grad - full.grad[sl]).abs().max() / full.grad.abs().max())
        # 3. whole-batch norm for the 3-D step rule
        u = torch.randn(global_batch, 3, 6, 6, 6)
        res["norm2"] = float(ctx.global_norm2((u[sl] ** 2).sum()))
        res["norm2_ref"] = float((u ** 2).sum())
        out_q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [4, 5])
def test_two_shards_reproduce_the_unsharded_scalars(global_batch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r["rank"])
    covered = []
    for r in results:
        covered += list(range(*r["slice"]))
        assert r["minmax"][0] == pytest.approx(r["minmax_ref"][0], abs=0)
        assert r["minmax"][1] == pytest.approx(r["minmax_ref"][1], abs=0)
        assert r["loss_sum"] == pytest.approx(r["loss_ref"], rel=2e-6)
        assert r["grad_err"] < 2e-6
        assert r["norm2"] == pytest.approx(r["norm2_ref"], rel=1e-6)
    assert covered == list(range(global_batch))


def test_shard_slice_partition():
    from advchain_b200.augmentor.sharding import shard_slice
    for n in (1, 7, 8, 16, 256):
        for w in (1, 2, 3, 8):
            idx = []
            for r in range(w):
                s = shard_slice(n, r, w)
                idx += list(range(s.start, s.stop))
            assert idx == list(range(n))
