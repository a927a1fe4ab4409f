"""torchrun check (N GPUs): the data-parallel step applies the MEAN of the per-replica gradients, replicas start identical
(rank-0 broadcast) and stay identical; prints one JSON line on rank 0.  SURVEY.md section 7 hard part 6."""
import faulthandler, json, os, sys
faulthandler.dump_traceback_later(45, exit=True)
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200 import parallel
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from sed_b200.train import DataParallelTrainer
from sed_b200.utils.common import WeightedBCE
import refmodels

rank, local_rank, world = parallel.init_process_group("nccl")
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
crit = WeightedBCE(recall_factor=5, multi_frame=True)
torch.manual_seed(100 + rank)                         # deliberately different initial weights per rank
model = Cnn_AvgPooling(1, model_config=refmodels.MAIN_CFG).to(dev)
tr = DataParallelTrainer(model, crit, lr=1e-3, graph=bool(int(os.environ.get("GRAPH", "1"))))
p0 = tr.flat.param.clone()
gathered = [torch.empty_like(p0) for _ in range(world)]
dist.all_gather(gathered, p0)
same_start = all(torch.equal(gathered[0], g) for g in gathered)
print(f"[{rank}] start ok", file=sys.stderr, flush=True)

def shard(r, it):
    g = torch.Generator().manual_seed(1000 * it + r)
    return (torch.randn(16, 1, 30, 64, generator=g) * 1.5).to(dev), (torch.rand(16, 30, 1, generator=g) > 0.8).float().to(dev)

# reference for step 1: every rank computes the gradients of ALL shards from the common initial weights
ref_model = Cnn_AvgPooling(1, model_config=refmodels.MAIN_CFG).to(dev)
ref_model.load_state_dict(model.state_dict())
mean_grad = torch.zeros_like(p0)
for r in range(world):
    ref_model.load_state_dict(model.state_dict())     # same running statistics for every shard
    ref_model.zero_grad()
    ref_model.train()
    x, y = shard(r, 0)
    crit(ref_model(x), y).backward()
    mean_grad += torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()]) / world
print(f"[{rank}] reference grads done", file=sys.stderr, flush=True)
x, y = shard(rank, 0)
loss = tr.step(x, y)
torch.cuda.synchronize()
print(f"[{rank}] step 1 done", file=sys.stderr, flush=True)
rel = float((tr.flat.grad / world - mean_grad).norm() / mean_grad.norm())      # the bucket holds the all-reduced SUM
for it in range(1, 4):
    x, y = shard(rank, it)
    loss = tr.step(x, y)
torch.cuda.synchronize()      # replayed NCCL kernels and eager collectives of one communicator must not overlap
dist.all_gather(gathered, tr.flat.param)
same_end = all(torch.equal(gathered[0], g) for g in gathered)
moved = float((tr.flat.param - p0).abs().max())
if rank == 0:
    print(json.dumps({"world": world, "same_start": same_start, "allreduced_grad_vs_mean_of_replicas_rel": rel,
                      "replicas_identical_after_4_steps": same_end, "max_param_move": moved, "loss": float(loss)}))
tr.close()
torch.cuda.synchronize()
dist.destroy_process_group()
