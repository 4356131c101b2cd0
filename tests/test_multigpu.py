"""Multi-GPU path on real GPUs (skipped with fewer than two): the batch sharded over NCCL ranks, the result gathered (a) by
ncclAllGather after the kernel and (b) by the fused gather -- the kernel storing v̇ into every GPU's peer-mapped array
(rbd_dynamics_gather).  Both must be bit-identical to the single-GPU evaluation of the whole batch (BASELINE config 5's path)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Bl, out_dir):
    import torch
    import torch.distributed as dist
    import rigidbodydynamics.jl_b200 as rbd
    from rigidbodydynamics.jl_b200.sharding import GatheredResult, dynamics_gather_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        mech = rbd.load_model("atlas", floating=True)
        B = world * Bl
        rng = np.random.default_rng(5)                     # every rank generates the same global batch, keeps its shard
        full = rbd.MechanismState(mech, B, torch.float32)
        rbd.rand_(full, rng)
        tau = torch.from_numpy(rng.random((full.nv, B))).float().cuda()
        lo = rank * Bl
        st = rbd.MechanismState(mech, Bl, torch.float32)
        st.q.copy_(full.q[:, lo:lo + Bl]); st.v.copy_(full.v[:, lo:lo + Bl])
        tl = tau[:, lo:lo + Bl].contiguous()
        res = rbd.DynamicsResult(mech, Bl, torch.float32)
        rbd.dynamics_(res, st, tl, want_qd=False)
        nccl = torch.empty((world, full.nv, Bl), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(nccl, res.vd)
        g = GatheredResult(full.nv, Bl, torch.float32)
        g.tensor.fill_(float("nan"))
        g.barrier(); torch.cuda.synchronize(); dist.barrier()
        dynamics_gather_(g, st, tl)
        spec = rbd.launch_info().specialised
        g.barrier(); torch.cuda.synchronize(); dist.barrier()
        whole = rbd.DynamicsResult(mech, B, torch.float32)
        rbd.dynamics_(whole, full, tau, want_qd=False)       # single-GPU evaluation of everything
        torch.cuda.synchronize()
        ok_nccl = torch.equal(nccl.permute(1, 0, 2).reshape(full.nv, B), whole.vd)
        ok_fused = torch.equal(g.tensor, whole.vd)
        with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
            f.write(f"{int(ok_nccl)} {int(ok_fused)} {int(spec)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("Bl", [1 << 15, 5000])
def test_sharded_dynamics_with_nccl_and_fused_gather(built, tmp_path, Bl):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    mp.spawn(_worker, args=(world, _free_port(), Bl, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok_nccl, ok_fused, spec = (int(x) for x in open(tmp_path / f"r{r}.txt").read().split())
        assert ok_nccl == 1, f"rank {r}: NCCL-gathered result differs from the single-GPU evaluation"
        assert ok_fused == 1, f"rank {r}: fused gather differs from the single-GPU evaluation"
