"""Data-parallel training for the reference's loop (SURVEY.md sections 8e/8f-4; reference train.py:77-132).

What the reference does per iteration (train.py:96-105): ``model.train(); out = model(x); loss = criterion(out, y);
zero_grad; backward; Adam(amsgrad=True).step(); loss.item()``.  This module keeps those semantics and makes them
native and multi-GPU:

* ``DataParallelTrainer.step`` runs the whole iteration on the kernels of libsedb.so: train-mode forward with batch
  statistics, WeightedBCE, backward (``sedb_cnn_train_forward`` / ``sedb_bce_with_logits`` / ``sedb_cnn_train_backward``),
  ONE NCCL all-reduce of the flat gradient bucket (2.33 MB for the main.py model), and the fused Adam-amsgrad update with
  the 1/world_size of the gradient mean folded in.  With ``graph=True`` the iteration -- all-reduce included -- is captured
  once into a CUDA graph and replayed; the step count and learning rate live in device memory for that.  (NCCL wants
  the operations of one communicator serialised: after graph-replayed steps, synchronise the stream before issuing an
  eager collective on the same process group.)
* parameters and gradients live in ONE flat float32 buffer each (the module's tensors are views);
* replicas start identical: rank 0's parameters and BatchNorm buffers are broadcast at construction (the reference's
  main.py does not seed model construction); BatchNorm statistics then stay per replica, as in the reference (no SyncBN);
* the loss is returned as a device tensor: no ``loss.item()`` sync per step (train.py:105 forces one);
* ``state_dict`` / ``load_state_dict`` expose the optimizer state in ``torch.optim.Adam``'s layout, so the checkpoint
  format of train.py:123-128 (``{'iterations', 'model', 'optimizer'}``) is kept.

``train(...)`` is the drop-in for the reference's function of the same name: LR decay every 200 iterations, logging,
rank-0-only evaluation through the native inference path, and the reference's checkpoint files.

Any model/criterion other than ``Cnn_AvgPooling`` + ``WeightedBCE(multi_frame=True)`` (e.g. M5) takes the generic path:
autograd forward/backward, same bucket, same all-reduce, same fused update.
"""
from __future__ import annotations

import ctypes
import os
import time

import torch
import torch.distributed as dist

from . import _ext


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


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
            p.grad = self.grad[off:off + k].view_as(p)             # gradients are written straight into the bucket
            off += k
        self.params = params
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat: torch.Tensor):
    """The single gradient collective of a step (no-op without a process group)."""
    if _world() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def broadcast_module_(module: torch.nn.Module, flat_param: torch.Tensor | None = None, src: int = 0):
    """Make every rank start from rank `src`'s parameters and buffers (what DistributedDataParallel does at wrap time)."""
    if _world() <= 1:
        return
    if flat_param is not None:
        dist.broadcast(flat_param, src=src)
    else:
        for p in module.parameters():
            dist.broadcast(p.data, src=src)
    for b in module.buffers():
        dist.broadcast(b, src=src)


class DataParallelTrainer:
    """One process per GPU; every rank holds a replica and a shard of the global batch."""

    def __init__(self, model, criterion, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, graph=False):
        self.model, self.criterion = model, criterion
        self.betas, self.eps, self.weight_decay = betas, float(eps), float(weight_decay)
        self.flat = FlatBuffers(model)
        if not self.flat.param.is_cuda:
            raise RuntimeError("DataParallelTrainer needs the model on a CUDA device (the fused update has no CPU "
                               "fallback)")
        self.world = _world()
        broadcast_module_(model, self.flat.param)
        dev = self.flat.param.device
        z = lambda: torch.zeros_like(self.flat.param)      # noqa: E731
        self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq = z(), z(), z()
        # device-resident optimizer scalars: [steps done, lr] and two floats of scratch (bias-corrected step size, ...)
        self._state = torch.tensor([0.0, float(lr)], dtype=torch.float32, device=dev)
        self._hyper = torch.zeros(2, dtype=torch.float32, device=dev)
        self._lr = float(lr)
        self.step_count = 0
        self.graph = bool(graph)
        self._graphs = {}          # (shape of x, shape of target) -> (CUDAGraph, static x, static target, static loss)
        self._loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._dlogits = None

    # ------------------------------------------------------------------ learning rate (train.py:108-110)
    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = float(value)
        self._state[1] = self._lr

    def decay_lr(self, factor=0.997):
        """train.py:108-110: lr *= 0.997 every 200 iterations."""
        self.lr = self._lr * factor

    # ------------------------------------------------------------------ one iteration
    def _native_ok(self, x, target):
        from .models.spectogram_models import Cnn_AvgPooling
        from .utils.common import WeightedBCE
        return (isinstance(self.model, Cnn_AvgPooling) and self.model.native_training
                and isinstance(self.criterion, WeightedBCE) and self.criterion.multi_frame
                and x.is_cuda and x.dim() == 4 and target.dim() == 3 and x.shape[0] > 0)

    def _step_native(self, x, target):
        """forward, loss, backward on the native kernels; gradients land in the flat bucket (no autograd graph)."""
        lib = _ext.load()
        m = self.model
        m.train()
        x = x.to(torch.float32).contiguous()
        target = target.to(torch.float32).contiguous()
        token = m._train_forward_native(x)
        out = m._train_out
        if self._dlogits is None or self._dlogits.shape != out.shape:
            self._dlogits = torch.empty_like(out)
        B, F_out, K = out.shape
        with torch.cuda.device(x.device):
            _ext.check(lib.sedb_bce_with_logits(_p(out), _p(target), B, F_out, target.shape[1], K,
                                                float(self.criterion.recall_factor), 1.0, _p(self._loss),
                                                _p(self._dlogits), _ext.stream_ptr()))
            h = m._train_handle(x.device)
            ws = m._native.workspace(x.device, ("train", B, x.shape[2]), 0)
            from .models._native import aligned_ptr
            ws_ptr, ws_bytes = aligned_ptr(ws)
            tensors = m._native_tensors()
            grads = [p.grad for p in m._train_params()]
            _ext.check(lib.sedb_cnn_train_backward(h, m._tensor_array(tensors), len(tensors), _p(x), _p(self._dlogits), B,
                                                   x.shape[2], m._tensor_array(grads), len(grads), ws_ptr, ws_bytes,
                                                   _ext.stream_ptr()))
        del token
        return self._loss

    def _step_generic(self, x, target):
        self.model.train()
        self.flat.zero_grad()
        loss = self.criterion(self.model(x), target)
        loss.backward()
        self._loss.copy_(loss.detach().reshape(1))
        return self._loss

    def _iteration(self, x, target):
        loss = self._step_native(x, target) if self._native_ok(x, target) else self._step_generic(x, target)
        allreduce_sum_(self.flat.grad)
        self._apply_update_dev()
        return loss

    def step(self, x, target):
        """forward + loss + backward + one all-reduce + fused Adam-amsgrad; returns the (local) loss as a device tensor
        (valid until the next step)."""
        self.step_count += 1
        if not self.graph:
            loss = self._iteration(x, target)
        else:
            key = (tuple(x.shape), tuple(target.shape), x.dtype, target.dtype)
            entry = self._graphs.get(key)
            if entry is None:
                entry = self._capture(x, target)
                self._graphs[key] = entry
            g, sx, st = entry
            sx.copy_(x)
            st.copy_(target)
            g.replay()
            loss = self._loss
        self._mark_updated()
        return loss

    def _capture(self, x, target):
        """Warm up on a side stream (plans, workspaces, NCCL) with the optimizer state saved and restored, then capture."""
        sx, st = x.clone(), target.clone()
        saved = [t.clone() for t in (self.flat.param, self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq, self._state)]
        bufs = [b.clone() for b in self.model.buffers()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._iteration(sx, st)
        torch.cuda.current_stream().wait_stream(s)
        for t, v in zip((self.flat.param, self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq, self._state), saved):
            t.copy_(v)
        for b, v in zip(self.model.buffers(), bufs):
            b.copy_(v)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._iteration(sx, st)
        # the capture pass did not execute anything: state is still the restored one
        return g, sx, st

    def close(self):
        """Drop the captured graphs.  Call before ``dist.destroy_process_group()``: a live CUDA graph that holds NCCL
        kernels keeps the communicator busy and the teardown waits for it forever."""
        self._graphs.clear()

    # ------------------------------------------------------------------ optimizer
    def _apply_update_dev(self):
        with torch.cuda.device(self.flat.param.device):
            _ext.check(_ext.load().sedb_adam_amsgrad_step_dev(
                _p(self.flat.param), _p(self.flat.grad), _p(self.exp_avg), _p(self.exp_avg_sq), _p(self.max_exp_avg_sq),
                self.flat.numel, _p(self._state), _p(self._hyper), self.betas[0], self.betas[1], self.eps,
                self.weight_decay, 1.0 / self.world, _ext.stream_ptr()))

    def apply_update(self):
        """Fused Adam-amsgrad on the flat buffers (gradients already summed over the ranks)."""
        self.step_count += 1
        self._apply_update_dev()
        self._mark_updated()

    def _mark_updated(self):
        # in-place update of tensors the native inference handle may have packed: bump their version counters
        bump = torch.autograd.graph.increment_version
        for q in self.flat.params:
            bump(q)                                        # no kernel launch

    # ------------------------------------------------------------------ checkpoint (train.py:123-128)
    def state_dict(self):
        """Optimizer state in torch.optim.Adam(amsgrad=True)'s layout: loadable by the reference's optimizer."""
        state, off = {}, 0
        for i, p in enumerate(self.flat.params):
            k = p.numel()
            sl = slice(off, off + k)
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[sl].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[sl].view_as(p).clone(),
                        "max_exp_avg_sq": self.max_exp_avg_sq[sl].view_as(p).clone()}
            off += k
        group = {"lr": self._lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": True, "params": list(range(len(self.flat.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        group = sd["param_groups"][0]
        self.betas, self.eps = tuple(group["betas"]), float(group["eps"])
        self.weight_decay = float(group["weight_decay"])
        self.lr = group["lr"]
        off, steps = 0, 0
        for i, p in enumerate(self.flat.params):
            k = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                self.max_exp_avg_sq[off:off + k].copy_(st["max_exp_avg_sq"].reshape(-1))
                steps = int(float(st["step"]))
            off += k
        self.step_count = steps
        self._state[0] = float(steps)


# ---------------------------------------------------------------------------------------------- reference train()
def eval(model, dataloader, criterion, outputs_dir, iteration, device, limit_val_samples=None):     # noqa: A001
    """train.py:12-74 without the plotting: validation losses and the 21-threshold recall / precision sets and AP per
    clip, through the native inference path (model.eval())."""
    from .utils.metric_utils import calculate_metrics
    losses, recal_sets, precision_sets, APs = [], [], [], []
    val_sampler = dataloader.dataset.get_validation_sampler(max_validate_num=limit_val_samples)
    was_training = model.training
    model.eval()
    for (input, target, file_name) in val_sampler:                                                  # noqa: A002
        with torch.no_grad():
            output = model(input.to(device).float()).cpu()
        losses.append(float(criterion(output, target.float())))
        if input.dim() == 4:
            output, target = output[0], target[0]
        else:
            target = target.reshape(-1, 1)
        r, p, ap = calculate_metrics(torch.sigmoid(output).numpy(), target.numpy())
        recal_sets.append(r)
        precision_sets.append(p)
        APs.append(ap)
    model.train(was_training)
    return losses, recal_sets, precision_sets, APs


def train(model, data_loader, criterion, num_steps, lr, log_freq, outputs_dir, device, graph=True, log=print):
    """Drop-in for the reference's ``train`` (train.py:77-132): same arguments, LR decay, logging cadence and checkpoint
    files.  Under torchrun every rank trains on its own loader shard; only rank 0 evaluates, logs and writes."""
    lr_decay_freq = 200
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if rank == 0:
        os.makedirs(os.path.join(outputs_dir, 'checkpoints'), exist_ok=True)
    model = model.to(device)
    trainer = DataParallelTrainer(model, criterion, lr=lr, betas=(0.9, 0.999), eps=1e-08, weight_decay=0.0, graph=graph)
    iterations, epoch = 0, 0
    t0 = time.time()
    history = {"train_loss": [], "val": []}
    pending = []                                   # device loss tensors not yet read back (no per-step sync)
    while iterations < num_steps:
        for (batch_features, event_labels) in data_loader:
            loss = trainer.step(batch_features.to(device, non_blocking=True).float(),
                                event_labels.to(device, non_blocking=True).float())
            pending.append(loss.clone())
            iterations += 1
            if iterations % lr_decay_freq == 0:
                trainer.decay_lr(0.997)
            if iterations % log_freq == 0:
                history["train_loss"] += [float(v) for v in torch.cat(pending).cpu()]
                pending = []
                if rank == 0:
                    im_sec = iterations * getattr(data_loader, "batch_size", 1) / (time.time() - t0)
                    log(f"epoch: {epoch}, step: {iterations}, loss: {history['train_loss'][-1]:.2f}, "
                        f"im/sec: {im_sec:.1f}, lr: {trainer.lr:.8f}")
                    if hasattr(data_loader.dataset, "get_validation_sampler"):
                        history["val"].append((iterations,) + tuple(eval(model, data_loader, criterion, outputs_dir,
                                                                         iteration=iterations, device=device,
                                                                         limit_val_samples=3)))
                    checkpoint = {'iterations': iterations, 'model': model.state_dict(),
                                  'optimizer': trainer.state_dict()}
                    torch.save(checkpoint, os.path.join(outputs_dir, 'checkpoints', f"iteration_{iterations}.pth"))
                if _world() > 1:
                    dist.barrier()
            if iterations == num_steps:
                break
        epoch += 1
    if pending:
        history["train_loss"] += [float(v) for v in torch.cat(pending).cpu()]
    return trainer, history
