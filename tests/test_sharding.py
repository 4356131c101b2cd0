"""N > 1 path on CPU: world_size-2 gloo processes shard a batch, evaluate their shards independently (no data-path
collective) and all-gather the result; the gathered array must equal the single-process evaluation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rigidbodydynamics.jl_b200 as rbd
from rigidbodydynamics.jl_b200.sharding import gather_columns, shard_bounds
from tests.util import rand_inputs


def test_shard_bounds_cover_batch_exactly():
    for B in (0, 1, 7, 64, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import Oracle                     # the CPU stand-in for the per-rank GPU evaluation
        mech = rbd.load_model("iiwa14")
        q, v, tau, _, _ = rand_inputs(mech, B, 123)   # every rank generates the same global batch, keeps its shard
        lo, hi = shard_bounds(B, world, rank)
        local = Oracle(mech.flatten()).dynamics(q[:, lo:hi], v[:, lo:hi], tau[:, lo:hi])
        full = gather_columns(torch.from_numpy(local), B)
        if rank == 0:
            np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 37])
def test_two_rank_shard_and_gather(tmp_path, B):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    from oracle import Oracle
    mech = rbd.load_model("iiwa14")
    q, v, tau, _, _ = rand_inputs(mech, B, 123)
    ref = Oracle(mech.flatten()).dynamics(q, v, tau)
    assert np.array_equal(np.load(tmp_path / "gathered.npy"), ref)
