"""one data-parallel training step per GPU on 64 ten-second crops -- fused log-mel ->
Cnn_AvgPooling (train mode) -> WeightedBCE -> backward -> one all-reduce -> fused Adam-amsgrad -- with the time split.
Run on 1 GPU directly or under torchrun for N GPUs.  Development measurement (not the headline bench)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200 import parallel
from sed_b200.dataset.spectogram import preprocess as P
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from sed_b200.train import DataParallelTrainer, allreduce_sum_
from sed_b200.utils.common import WeightedBCE

rank, local_rank, world = parallel.init_process_group()
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
torch.manual_seed(0)
model = Cnn_AvgPooling(1, model_config=[(32, 2), (64, 2), (128, 2), (128, 1)]).to(dev)
tr = DataParallelTrainer(model, WeightedBCE(recall_factor=5, multi_frame=True), lr=1e-6)
g = torch.Generator(device=dev).manual_seed(1234 + rank)
wave = (torch.randn(64, 480000, device=dev, generator=g) * 0.1).clamp_(-1, 1)
target = (torch.rand(64, 30, 1, device=dev, generator=g) > 0.8).float()
ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731
steps, warm = 20, 5
parts = {k: 0.0 for k in ("logmel", "forward_loss", "backward", "allreduce", "update")}
tot = 0.0
for it in range(warm + steps):
    e = [ev() for _ in range(6)]
    parallel.barrier()
    e[0].record()
    x = P.waveform_to_log_mel(wave)[:, None, :30]                      # (64, 1, 30, 64): train_crop_size frames
    e[1].record()
    model.train()
    loss = tr.criterion(model(x), target)
    e[2].record()
    tr.flat.zero_grad()
    loss.backward()
    e[3].record()
    allreduce_sum_(tr.flat.grad)
    e[4].record()
    tr.apply_update()
    e[5].record()
    torch.cuda.synchronize()
    if it >= warm:
        for k, a, b in zip(parts, e[:-1], e[1:]):
            parts[k] += a.elapsed_time(b)
        tot += e[0].elapsed_time(e[5])
ms = parallel.max_over_ranks(tot / steps, dev)
if rank == 0:
    out = {"config": "64 crops x 10 s per GPU, Cnn_AvgPooling(32-64-128-128) train step", "n_gpus": world,
           "ms_per_step": ms, "steps_per_s": 1e3 / ms, "audio_hours_per_s": world * 64 * 10 / 3600 / (ms * 1e-3),
           "split_ms": {k: v / steps for k, v in parts.items()}, "loss": float(loss)}
    print(json.dumps(out))
