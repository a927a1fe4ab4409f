"""Drop-in for the feature-extraction functions of the reference's ``dataset/spectogram/preprocess.py``.

Same names, argument meaning and result types as the reference:

* ``MEL_FILTER_BANK_MATRIX``                 (preprocess.py:13-18)   (16385, 64) float32
* ``multichannel_stft(signal)``              (preprocess.py:21-36)   (samples, C) -> (C, T, 16385) complex64
* ``multichannel_complex_to_log_mel(spec)``  (preprocess.py:39-45)   complex (C,T,16385) | (T,16385) -> float32
* ``calculate_scalar_of_tensor(x)``          (preprocess.py:48-57)

plus the fused fast path the reference does not have:

* ``waveform_to_log_mel(wave, mean=None, std=None)``: CUDA float32 ``[B, samples]`` -> ``[B, T, 64]`` without
  materialising the complex STFT (== ``multichannel_complex_to_log_mel(multichannel_stft(.))`` per clip).

All arithmetic runs in libsedb.so's sm_100a kernels.  NumPy inputs are copied to the current CUDA device and
the result copied back; CUDA tensors stay on the device.  There is no CPU fallback: inside a DataLoader worker
(no CUDA context) these functions raise.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import spectogram_configs as cfg
from ... import _ext


def _mel_matrix() -> np.ndarray:
    lib = _ext.load()
    _ext.check(lib.sedb_check_config(cfg.working_sample_rate, cfg.frame_size, cfg.hop_size, cfg.NFFT, cfg.mel_bins,
                                     float(cfg.mel_min_freq), float(cfg.mel_max_freq)))
    out = np.empty((cfg.NFFT // 2 + 1, cfg.mel_bins), dtype=np.float32)
    _ext.check(lib.sedb_mel_filterbank(ctypes.c_void_p(out.ctypes.data)))
    return out


MEL_FILTER_BANK_MATRIX = _mel_matrix()
NUM_BINS = cfg.NFFT // 2 + 1


def num_frames(n_samples: int) -> int:
    return 1 + int(n_samples) // cfg.hop_size


def _as_cuda_f32(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.cuda()
        return x.to(torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()


def _row_stride(w) -> int:
    # torch reports an arbitrary stride for a size-1 leading dimension
    return w.stride(0) if w.shape[0] > 1 else max(w.stride(0), w.shape[1])


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _norm_tensor(mean, std):
    if mean is None and std is None:
        return None
    if mean is None or std is None:
        raise ValueError("mean and std must be given together")
    m = _as_cuda_f32(mean).reshape(-1)
    s = _as_cuda_f32(std).reshape(-1)
    if m.numel() != cfg.mel_bins or s.numel() != cfg.mel_bins:
        raise ValueError(f"mean/std must have {cfg.mel_bins} entries (per mel bin)")
    return torch.cat([m, s]).contiguous()


def waveform_to_log_mel(wave, mean=None, std=None) -> torch.Tensor:
    """Fused log-mel of a batch of mono clips: ``[B, samples]`` (or ``[samples]``) -> ``[B, T, 64]`` float32 CUDA."""
    w = _as_cuda_f32(wave)
    squeeze = w.dim() == 1
    if squeeze:
        w = w[None]
    if w.dim() != 2:
        raise ValueError("waveform_to_log_mel expects [B, samples] or [samples]")
    B, n = w.shape
    norm = _norm_tensor(mean, std)
    out = torch.empty((B, num_frames(n), cfg.mel_bins), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _ext.check(_ext.load().sedb_logmel_f32(_ext.context(), _ptr(w), B, n, _row_stride(w), _ptr(norm), _ptr(out),
                                               _ext.stream_ptr()))
    return out[0] if squeeze else out


def multichannel_stft(multichannel_signal):
    """(samples, channels) -> (channels, T, 16385) complex64; NumPy in -> NumPy out, CUDA tensor in -> CUDA out."""
    is_np = not isinstance(multichannel_signal, torch.Tensor)
    if multichannel_signal.ndim != 2:
        raise ValueError("multichannel_stft expects (samples, channels)")
    w = _as_cuda_f32(multichannel_signal).t().contiguous()          # [C, samples]
    C, n = w.shape
    spec = torch.empty((C, num_frames(n), NUM_BINS), dtype=torch.complex64, device=w.device)
    with torch.cuda.device(w.device):
        _ext.check(_ext.load().sedb_stft_c64(_ext.context(), _ptr(w), C, n, _row_stride(w), _ptr(spec),
                                             _ext.stream_ptr()))
    return spec.cpu().numpy() if is_np else spec


def multichannel_complex_to_log_mel(multichannel_complex_spectogram, mean=None, std=None):
    """complex (C,T,16385) or (T,16385) -> float32 log-mel of the same leading shape with 64 mel bins."""
    x = multichannel_complex_spectogram
    is_np = not isinstance(x, torch.Tensor)
    if is_np:
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.complex64))
    if not x.is_cuda:
        x = x.cuda()
    x = x.to(torch.complex64).contiguous()
    if x.shape[-1] != NUM_BINS:
        raise ValueError(f"last dimension must be {NUM_BINS} rFFT bins")
    rows = x.numel() // NUM_BINS
    norm = _norm_tensor(mean, std)
    out = torch.empty(x.shape[:-1] + (cfg.mel_bins,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _ext.check(_ext.load().sedb_power_mel_db_f32(_ext.context(), _ptr(x), rows, _ptr(norm), _ptr(out),
                                                     _ext.stream_ptr()))
    return out.cpu().numpy() if is_np else out


def calculate_scalar_of_tensor(x):
    """Per-feature mean/std over all leading axes (reference preprocess.py:48-57): 2-D -> axis 0, 3-D -> (0,1)."""
    if isinstance(x, torch.Tensor):
        dims = 0 if x.ndim == 2 else (0, 1)
        return x.mean(dim=dims), x.std(dim=dims, unbiased=False)
    axis = 0 if x.ndim == 2 else (0, 1)
    return np.mean(x, axis=axis), np.std(x, axis=axis)
