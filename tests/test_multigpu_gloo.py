"""World-size-2 test of the N>1 host path on CPU (gloo): contiguous clip shards, no data-path collective, results
gathered on rank 0, timing reduced with max over ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_clips, ret):
    sys.path.insert(0, ROOT)
    import sed_b200  # noqa: F401
    from sed_b200 import parallel
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    a, b = parallel.shard_range(n_clips, rank, world)
    # stand-in for the per-clip hot path: every clip maps to a deterministic (frames x classes) block
    local = torch.stack([torch.full((4, 1), float(c)) for c in range(a, b)]) if b > a else torch.zeros(0, 4, 1)
    parallel.barrier()
    t = parallel.max_over_ranks(0.5 + rank)
    total = parallel.sum_over_ranks(b - a)
    full = parallel.gather_sharded(local, n_clips)
    if rank == 0:
        ret["t"], ret["total"], ret["full"] = t, total, full.clone()
    else:
        assert full is None
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    n_clips, world = 7, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29631, n_clips, ret), nprocs=world, join=True)
    assert ret["t"] == 1.5 and ret["total"] == n_clips
    full = ret["full"]
    assert full.shape == (n_clips, 4, 1)
    assert torch.equal(full[:, 0, 0], torch.arange(n_clips, dtype=torch.float32))


def _grad_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import sed_b200  # noqa: F401
    from sed_b200 import parallel
    from sed_b200.train import FlatBuffers, allreduce_sum_
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    parallel.init_process_group("gloo")
    torch.manual_seed(0)                                   # same replica on every rank
    net = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
    flat = FlatBuffers(net)
    x = torch.full((3, 8), float(rank + 1))
    net(x).sum().backward()                                # gradients land in the flat bucket
    local = flat.grad.clone()
    allreduce_sum_(flat.grad)                              # the ONE collective of a step
    ret[rank] = (local, flat.grad.clone(), flat.numel)
    dist.destroy_process_group()


def test_two_rank_single_bucket_gradient_allreduce():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, 29641, ret), nprocs=2, join=True)
    (l0, s0, n0), (l1, s1, n1) = ret[0], ret[1]
    assert n0 == n1 == 8 * 4 + 4 + 4 + 1
    assert torch.allclose(s0, l0 + l1) and torch.equal(s0, s1)
    assert not torch.equal(l0, l1)
