"""Drop-in for the reference's ``models/spectogram_models.py`` (``Cnn_AvgPooling``, ``ConvBlock``, ``interpolate``).

Same constructor arguments, attributes (``conv_blocks``, ``event_fc``, ``num_pools``, ``model_config``), parameter
initialisation and ``state_dict`` keys as the reference (spectogram_models.py:128-230), so checkpoints written by the
reference's ``train.py:123-128`` load unchanged and ``main.py:35`` / ``infer.py:21`` construct it the same way.

* eval mode (``model.eval()``; what ``train.py:21-24`` and ``infer.py:32-33`` run): ``forward``/``logits`` execute the
  hand-written sm_100a kernels of libsedb.so (implicit-GEMM tcgen05 convolutions with folded BatchNorm, fused
  head).  CUDA only -- a CPU tensor raises, there is no fallback.
* train mode on a CUDA tensor (``train.py:96-103``): ``forward`` is ONE autograd node backed by the native training
  kernels of libsedb.so -- batch-statistics BatchNorm forward (running statistics updated in place), and in
  ``backward`` the BN / ReLU / pooling / head gradients plus tcgen05 data- and weight-gradient convolutions -- so the
  reference loop ``loss = criterion(model(x), y); loss.backward(); optimizer.step()`` runs on them unchanged.  The input
  is treated as data (no gradient w.r.t. ``x``).  On CPU tensors, or with ``model.native_training = False``, train mode
  falls back to the differentiable torch expression of the same network (used by the CPU tests as the comparison).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..dataset.spectogram.spectogram_configs import audio_channels, working_sample_rate, mel_bins, hop_size, classes_num
from ..utils.common import count_parameters, human_format
from .. import _ext
from ._native import NativeHandle, _ptr, aligned_ptr

DEFAULT_CHANNEL_AND_POOL = [(64, 2), (128, 2), (256, 2), (512, 1)]


def interpolate(x, ratio):
    """Repeat every time step ``ratio`` times: (B, T, C) -> (B, T*ratio, C)  (reference spectogram_models.py:9-22)."""
    b, t, c = x.shape
    return x.unsqueeze(2).expand(b, t, ratio, c).reshape(b, t * ratio, c)


def init_layer(layer, nonlinearity='leaky_relu'):
    nn.init.kaiming_uniform_(layer.weight, nonlinearity=nonlinearity)
    if getattr(layer, 'bias', None) is not None:
        layer.bias.data.fill_(0.)


def init_bn(bn):
    bn.bias.data.fill_(0.)
    bn.running_mean.data.fill_(0.)
    bn.weight.data.fill_(1.)
    bn.running_var.data.fill_(1.)


class ConvBlock(nn.Module):
    """conv3x3-BN-ReLU x2 + AvgPool(pool_size); parameter container for the native kernels."""

    def __init__(self, in_channels, out_channels, pool_size=2):
        super().__init__()
        self.pool_size = pool_size
        conv = dict(kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.conv1 = nn.Conv2d(in_channels, out_channels, **conv)
        self.conv2 = nn.Conv2d(out_channels, out_channels, **conv)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.init_weights()

    def init_weights(self):
        for conv in (self.conv1, self.conv2):
            init_layer(conv)
        for bn in (self.bn1, self.bn2):
            init_bn(bn)

    def native_tensors(self):
        """Tensor order expected by sedb_cnn_load (include/sedb.h)."""
        out = [self.conv1.weight, self.conv2.weight]
        for bn in (self.bn1, self.bn2):
            out += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        return out

    def forward(self, input):
        # differentiable expression of the block (training); inference goes through Cnn_AvgPooling's native path
        x = F.relu(self.bn1(self.conv1(input)))
        x = F.relu(self.bn2(self.conv2(x)))
        return F.avg_pool2d(x, kernel_size=self.pool_size)


class _NativeTrainFunction(torch.autograd.Function):
    """forward: sedb_cnn_train_forward; backward: sedb_cnn_train_backward (include/sedb.h).  The activations live in
    the module's training workspace between the two calls, as in one iteration of train.py:96-103."""

    @staticmethod
    def forward(ctx, module, x, *params):
        ctx.module, ctx.x = module, x
        ctx.token = module._train_forward_native(x)
        ctx.mark_non_differentiable()
        return module._train_out

    @staticmethod
    def backward(ctx, dlogits):
        grads = ctx.module._train_backward_native(ctx.x, dlogits.contiguous(), ctx.token)
        return (None, None) + tuple(grads)


class Cnn_AvgPooling(nn.Module):
    native_training = True       # train-mode forward/backward of CUDA tensors on the native kernels

    def __init__(self, classes_num, model_config=DEFAULT_CHANNEL_AND_POOL):
        super().__init__()
        self.model_config = model_config
        self.classes_num = classes_num
        # the reference counts the first block as one pool whatever its size (spectogram_models.py:167-172)
        self.num_pools = 1 + sum(1 for (_, pool) in model_config[1:] if pool == 2)
        blocks, in_ch = [], audio_channels
        for (out_ch, pool) in model_config:
            blocks.append(ConvBlock(in_channels=in_ch, out_channels=out_ch, pool_size=pool))
            in_ch = out_ch
        self.conv_blocks = nn.Sequential(*blocks)
        self.event_fc = nn.Linear(model_config[-1][0], classes_num, bias=True)
        self.init_weights()
        self._native = None

    def init_weights(self):
        init_layer(self.event_fc)

    # ------------------------------------------------------------------ native inference
    def _native_tensors(self):
        ts = []
        for blk in self.conv_blocks:
            ts += blk.native_tensors()
        return ts + [self.event_fc.weight, self.event_fc.bias]

    def _handle_init(self):
        lib = _ext.load()
        if self._native is None:
            channels = (ctypes.c_int * len(self.model_config))(*[int(c) for c, _ in self.model_config])
            pools = (ctypes.c_int * len(self.model_config))(*[int(p) for _, p in self.model_config])
            n, classes = len(self.model_config), int(self.classes_num)

            def create(out):
                _ext.check(lib.sedb_cnn_create(_ext.context(), channels, pools, n, classes, out))

            def load(h, arr, cnt):
                _ext.check(lib.sedb_cnn_load(h, arr, cnt, _ext.stream_ptr()))

            self._native = NativeHandle(create, lib.sedb_cnn_destroy, load, lib.sedb_cnn_workspace_invalidate)

    def _handle(self, device):
        self._handle_init()
        return self._native.get(device, self._native_tensors())

    def _forward_native(self, x, want_probs):
        if not x.is_cuda:
            raise RuntimeError("Cnn_AvgPooling inference runs on CUDA (sm_100a) only; got a CPU tensor and there is "
                               "no CPU fallback")
        if x.dim() != 4 or x.shape[1] != audio_channels or x.shape[3] != mel_bins:
            raise ValueError(f"expected input (batch, {audio_channels}, time_steps, {mel_bins}), got {tuple(x.shape)}")
        lib = _ext.load()
        x = x.to(torch.float32).contiguous()
        B, _, T, _ = x.shape
        with torch.cuda.device(x.device):
            h = self._handle(x.device)
            out_frames = lib.sedb_cnn_out_frames(h, T)
            if out_frames <= 0:
                raise ValueError(f"input of {T} frames is too short for this model's pooling")
            out = torch.empty((B, out_frames, self.classes_num), dtype=torch.float32, device=x.device)
            if B == 0:
                return out
            need = lib.sedb_cnn_workspace_bytes(h, B, T)
            ws = self._native.workspace(x.device, (B, T), need)
            ws_ptr, ws_bytes = aligned_ptr(ws)
            _ext.check(lib.sedb_cnn_forward(h, _ptr(x), B, T, None if want_probs else _ptr(out),
                                            _ptr(out) if want_probs else None, ws_ptr, ws_bytes, _ext.stream_ptr()))
        return out

    # ------------------------------------------------------------------ native training step
    def _train_params(self):
        """module.parameters() order: the order of the gradient list of sedb_cnn_train_backward."""
        ps = []
        for blk in self.conv_blocks:
            ps += [blk.conv1.weight, blk.conv2.weight, blk.bn1.weight, blk.bn1.bias, blk.bn2.weight, blk.bn2.bias]
        return ps + [self.event_fc.weight, self.event_fc.bias]

    def _train_handle(self, device):
        """The native handle WITHOUT (re)loading folded inference weights: the training entry points read the tensors."""
        self._handle_init()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._native._handles.get(idx)
        if entry is None:
            h = ctypes.c_void_p()
            self._native._create(ctypes.byref(h))
            entry = [h, None]
            self._native._handles[idx] = entry
        return entry[0]

    def _tensor_array(self, tensors):
        for t in tensors:
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise RuntimeError("native training needs contiguous float32 CUDA parameters and buffers")
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def _train_forward_native(self, x):
        lib = _ext.load()
        if x.dim() != 4 or x.shape[1] != audio_channels or x.shape[3] != mel_bins:
            raise ValueError(f"expected input (batch, {audio_channels}, time_steps, {mel_bins}), got {tuple(x.shape)}")
        B, _, T, _ = x.shape
        with torch.cuda.device(x.device):
            h = self._train_handle(x.device)
            out_frames = lib.sedb_cnn_out_frames(h, T)
            if out_frames <= 0 or B == 0:
                raise ValueError(f"native training needs a non-empty batch of at least {2 ** self.num_pools} frames")
            need = lib.sedb_cnn_train_workspace_bytes(h, B, T)
            if need == 0:
                raise _ext.SedbError(lib.sedb_last_error().decode())
            ws = self._native.workspace(x.device, ("train", B, T), need)
            ws_ptr, ws_bytes = aligned_ptr(ws)
            self._train_out = torch.empty((B, out_frames, self.classes_num), dtype=torch.float32, device=x.device)
            tensors = self._native_tensors()
            momentum = self.conv_blocks[0].bn1.momentum
            _ext.check(lib.sedb_cnn_train_forward(h, self._tensor_array(tensors), len(tensors), _ptr(x), B, T,
                                                  0.1 if momentum is None else float(momentum), _ptr(self._train_out),
                                                  ws_ptr, ws_bytes, _ext.stream_ptr()))
            # bookkeeping torch would have done: num_batches_tracked, and the version of the buffers written behind its back
            bns = [bn for blk in self.conv_blocks for bn in (blk.bn1, blk.bn2)]
            torch._foreach_add_([bn.num_batches_tracked for bn in bns], 1)
            bump = torch.autograd.graph.increment_version
            for bn in bns:
                bump(bn.running_mean)
                bump(bn.running_var)
        self._train_token = getattr(self, "_train_token", 0) + 1
        return self._train_token

    def _train_backward_native(self, x, dlogits, token):
        if token != self._train_token:
            raise RuntimeError("native training keeps the activations of ONE forward pass: call backward before the next "
                               "train-mode forward (train.py:96-103 does)")
        lib = _ext.load()
        B, _, T, _ = x.shape
        params = self._train_params()
        grads = [torch.empty_like(p) for p in params]
        with torch.cuda.device(x.device):
            h = self._train_handle(x.device)
            ws = self._native.workspace(x.device, ("train", B, T), 0)
            ws_ptr, ws_bytes = aligned_ptr(ws)
            tensors = self._native_tensors()
            _ext.check(lib.sedb_cnn_train_backward(h, self._tensor_array(tensors), len(tensors), _ptr(x), _ptr(dlogits), B,
                                                   T, self._tensor_array(grads), len(grads), ws_ptr, ws_bytes,
                                                   _ext.stream_ptr()))
        return grads

    def _use_native_training(self, x):
        return (self.native_training and x.is_cuda and torch.is_grad_enabled()
                and all(p.requires_grad for p in self._train_params()))

    def train(self, mode: bool = True):
        if self._native is not None and mode != self.training:
            self._native.mark_dirty()         # BatchNorm statistics may change behind the version counters
        return super().train(mode)

    # ------------------------------------------------------------------ reference interface
    def forward(self, x):
        """Input (batch, channels, time_steps, freq_bins) -> frame logits (batch, time_steps', classes)."""
        if not self.training:
            with torch.no_grad():
                return self._forward_native(x, want_probs=False)
        if self._use_native_training(x):
            if x.requires_grad:
                raise NotImplementedError("the native training path treats the input as data (no gradient w.r.t. x); set "
                                          "model.native_training = False for that")
            return _NativeTrainFunction.apply(self, x.detach().to(torch.float32).contiguous(), *self._train_params())
        x = self.conv_blocks(x)
        x = torch.mean(x, dim=3).transpose(1, 2)
        return interpolate(self.event_fc(x), 2 ** self.num_pools)

    def logits(self, x):
        """Frame probabilities (the reference names this ``logits``: sigmoid(forward(x)), spectogram_models.py:204-205)."""
        if not self.training:
            with torch.no_grad():
                return self._forward_native(x, want_probs=True)
        return torch.sigmoid(self.forward(x))

    def model_description(self):
        print("Model description")
        h, w, c = 60 * working_sample_rate // hop_size, mel_bins, audio_channels
        print(f"\tInput: (b, {c}, {h}, {w})")
        for (c, k) in self.model_config:
            h, w = h // k, w // k
            print(f"\tconv_block -> (b, {c}, {h}, {w})")
        print(f"\tmean(dim=3) -> (b, {c}, {h})")
        print(f"\ttranspose(1,2) -> (b, {h}, {c})")
        print(f"\tFC + sigmoid -> (b, {h}, {classes_num})")
        ratio = 2 ** self.num_pools
        print(f"\tinterpolate({ratio})-> (b, {h * ratio}, {classes_num})")
        print(f"\tModel has {h} outputs before interpolation, each stands for {ratio} frames or"
              f" {ratio * hop_size / working_sample_rate:.2f}s")
        print(f"\tModel has {human_format(count_parameters(self))} parameters")
