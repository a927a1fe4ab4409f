"""Data-parallel training step for the reference's loop (reference train.py:77-132).

What the reference does per iteration (train.py:96-105): ``model.train(); out = model(x); loss = criterion(out, y);
zero_grad; backward; Adam(amsgrad=True).step(); loss.item()``.  This module keeps those semantics and makes them
multi-GPU:

* parameters and gradients live in ONE flat float32 buffer each (the tensors of the module become views), so a step
  needs exactly one gradient all-reduce (NCCL over NVLink; 2.33 MB for the main.py model) -- no per-tensor buckets;
* the optimizer update is one fused kernel of libsedb.so (``sedb_adam_amsgrad_step``) over the flat buffers, with the
  1/world_size of the gradient mean folded in;
* the loss is returned as a device tensor: no ``loss.item()`` sync per step (train.py:105 forces one).

The backward pass itself is torch autograd (cuDNN) in this round; the forward uses the module's differentiable
train-mode expression.  BatchNorm statistics are per replica, as in the reference (it has no SyncBN).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _ext


class FlatBuffers:
    """Re-homes a module's parameters (and their gradients) into two flat contiguous float32 buffers."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.param = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.param[off:off + k].view_as(p)            # parameter now aliases the flat buffer
            p.grad = self.grad[off:off + k].view_as(p)             # autograd accumulates straight into the bucket
            off += k
        self.params = params
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()


def allreduce_sum_(flat: torch.Tensor):
    """The single gradient collective of a step (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


class DataParallelTrainer:
    """One process per GPU; every rank holds a replica and a shard of the global batch."""

    def __init__(self, model, criterion, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model, self.criterion = model, criterion
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.flat = FlatBuffers(model)
        if not self.flat.param.is_cuda:
            raise RuntimeError("DataParallelTrainer needs the model on a CUDA device (the fused update has no CPU "
                               "fallback)")
        z = lambda: torch.zeros_like(self.flat.param)      # noqa: E731
        self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq = z(), z(), z()
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def step(self, x, target):
        """forward + loss + backward + one all-reduce + fused Adam-amsgrad; returns the (local) loss tensor."""
        self.model.train()
        self.flat.zero_grad()
        loss = self.criterion(self.model(x), target)
        loss.backward()
        allreduce_sum_(self.flat.grad)
        self.apply_update()
        return loss.detach()

    def apply_update(self):
        """Fused Adam-amsgrad on the flat buffers (gradients already summed over the ranks)."""
        self.step_count += 1
        p = lambda t: ctypes.c_void_p(t.data_ptr())        # noqa: E731
        with torch.cuda.device(self.flat.param.device):
            _ext.check(_ext.load().sedb_adam_amsgrad_step(
                p(self.flat.param), p(self.flat.grad), p(self.exp_avg), p(self.exp_avg_sq), p(self.max_exp_avg_sq),
                self.flat.numel, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                1.0 / self.world, _ext.stream_ptr()))
        # in-place update of tensors the native inference handle may have packed: bump their version counters
        bump = getattr(torch.autograd.graph, "increment_version", None)
        for q in self.flat.params:
            if bump is not None:
                bump(q)                                    # no kernel launch
            else:
                q.data.add_(0)

    def decay_lr(self, factor=0.997):
        """train.py:108-110: lr *= 0.997 every 200 iterations."""
        self.lr *= factor
