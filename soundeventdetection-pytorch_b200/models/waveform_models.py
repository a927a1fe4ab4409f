"""Drop-in for the reference's ``models/waveform_models.py`` (``M5``: raw-waveform 1-D CNN, waveform_models.py:9-75).

Same constructor, ``state_dict`` keys (``conv_block{1..5}.{0,1,3,4}.*``, ``fc.*``) and default torch initialisation.
eval-mode ``forward`` on CUDA runs the sm_100a kernels of libsedb.so (strided k=79 front convolution, implicit-GEMM
tcgen05 k=3 convolutions with folded BatchNorm and fused MaxPool, mean+Linear head); train mode is expressed with
differentiable torch ops.  Returns logits (the reference leaves its sigmoid commented out, waveform_models.py:69).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import Sequential

from ..dataset.waveform.waveform_configs import frame_size, audio_channels
from ..utils.common import count_parameters, human_format
from .. import _ext
from ._native import NativeHandle, _ptr, aligned_ptr

# (in_channels, out_channels, max-pool after the block) for conv_block2..5: two k=3 convolutions each
_K3_BLOCKS = [(64, 64, True), (64, 64, True), (64, 128, True), (128, 256, False)]


def _conv_bn_relu(cin, cout, **conv_kwargs):
    return [nn.Conv1d(cin, cout, **conv_kwargs), nn.BatchNorm1d(cout), nn.ReLU()]


class M5(nn.Module):
    """Model of "Very deep convolutional neural networks for raw waveforms" as configured by the reference."""

    def __init__(self, classes_num):
        super().__init__()
        self.classes_num = classes_num
        self.conv_block1 = Sequential(*_conv_bn_relu(audio_channels, 64, kernel_size=79, stride=4, padding=39),
                                      nn.MaxPool1d(4, 4))
        for idx, (cin, cout, pool) in enumerate(_K3_BLOCKS, start=2):
            layers = _conv_bn_relu(cin, cout, kernel_size=3, stride=1, padding=1) \
                + _conv_bn_relu(cout, cout, kernel_size=3, stride=1, padding=1)
            if pool:
                layers.append(nn.MaxPool1d(4, 4))
            setattr(self, f"conv_block{idx}", Sequential(*layers))
        self.fc = nn.Linear(256, classes_num)
        self._native = None

    def _blocks(self):
        return [getattr(self, f"conv_block{i}") for i in range(1, 6)]

    def _native_tensors(self):
        """Order expected by sedb_m5_load: per Conv1d+BatchNorm1d pair w, b, gamma, beta, mean, var; then fc."""
        ts = []
        for blk in self._blocks():
            mods = list(blk)
            for i, m in enumerate(mods):
                if isinstance(m, nn.Conv1d):
                    bn = mods[i + 1]
                    ts += [m.weight, m.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        return ts + [self.fc.weight, self.fc.bias]

    def _handle(self, device):
        lib = _ext.load()
        if self._native is None:
            classes = int(self.classes_num)

            def create(out):
                _ext.check(lib.sedb_m5_create(_ext.context(), classes, out))

            def load(h, arr, cnt):
                _ext.check(lib.sedb_m5_load(h, arr, cnt, _ext.stream_ptr()))

            self._native = NativeHandle(create, lib.sedb_m5_destroy, load, lib.sedb_m5_workspace_invalidate)
        return self._native.get(device, self._native_tensors())

    def _forward_native(self, x):
        if not x.is_cuda:
            raise RuntimeError("M5 inference runs on CUDA (sm_100a) only; got a CPU tensor and there is no CPU fallback")
        if x.dim() != 3 or x.shape[1] != audio_channels or x.shape[2] != frame_size:
            raise ValueError(f"expected input (batch, {audio_channels}, {frame_size}), got {tuple(x.shape)}")
        lib = _ext.load()
        x = x.to(torch.float32).contiguous()
        n = x.shape[0]
        out = torch.empty((n, self.classes_num), dtype=torch.float32, device=x.device)
        if n == 0:
            return out
        with torch.cuda.device(x.device):
            h = self._handle(x.device)
            need = lib.sedb_m5_workspace_bytes(h, n)
            ws = self._native.workspace(x.device, n, need)
            ws_ptr, ws_bytes = aligned_ptr(ws)
            _ext.check(lib.sedb_m5_forward(h, _ptr(x), n, _ptr(out), ws_ptr, ws_bytes, _ext.stream_ptr()))
        return out

    def train(self, mode: bool = True):
        if self._native is not None and mode != self.training:
            self._native.mark_dirty()         # BatchNorm statistics may change behind the version counters
        return super().train(mode)

    def forward(self, x):
        # x: (b, c, frame_size) -> (b, classes) logits
        if not self.training:
            with torch.no_grad():
                return self._forward_native(x)
        for blk in self._blocks():
            x = blk(x)
        return self.fc(torch.mean(x, dim=2))

    def model_description(self):
        print("Waveform model:")
        print(f"\t- Model has {human_format(count_parameters(self))} parameters")
