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


def _bcast_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import sed_b200  # noqa: F401
    from sed_b200 import parallel
    from sed_b200.models.spectogram_models import Cnn_AvgPooling
    from sed_b200.train import FlatBuffers, broadcast_module_
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    parallel.init_process_group("gloo")
    torch.manual_seed(100 + rank)                          # the reference's main.py does not seed model construction
    net = Cnn_AvgPooling(1, model_config=[(16, 2), (16, 1)])
    with torch.no_grad():
        net.conv_blocks[0].bn1.running_mean.fill_(float(rank))
    flat = FlatBuffers(net)
    before = flat.param.clone()
    broadcast_module_(net, flat.param)                     # what DataParallelTrainer does at construction
    ret[rank] = (before, flat.param.clone(), net.conv_blocks[0].bn1.running_mean.clone(),
                 net.event_fc.weight.data_ptr() == flat.param[-17:-1].data_ptr())
    dist.destroy_process_group()


def test_two_rank_replicas_start_from_rank0_parameters_and_buffers():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bcast_worker, args=(2, 29651, ret), nprocs=2, join=True)
    (b0, a0, rm0, v0), (b1, a1, rm1, v1) = ret[0], ret[1]
    assert not torch.equal(b0, b1)                         # different initial weights ...
    assert torch.equal(a0, b0) and torch.equal(a1, b0)     # ... every rank ends up with rank 0's
    assert torch.equal(rm0, rm1) and float(rm1[0]) == 0.0  # buffers too
    assert v0 and v1                                       # the module's tensors are views of the flat buffer
