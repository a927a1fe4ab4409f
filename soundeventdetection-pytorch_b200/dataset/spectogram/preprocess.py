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


def _is_int16(x) -> bool:
    return (isinstance(x, torch.Tensor) and x.dtype == torch.int16) or (isinstance(x, np.ndarray) and x.dtype == np.int16)


def pcm16_to_log_mel(pcm, mean=None, std=None) -> torch.Tensor:
    """Fused log-mel straight from 16-bit PCM (the WAV data chunk): ``[B, samples]`` or ``[B, samples, channels]`` int16
    -> ``[B, T, 64]`` float32 CUDA.  The kernel forms the mono mix ``mean_ch(s / 32768)`` while loading, which is what
    the reference's ``read_multichannel_audio`` produces for ``audio_channels == 1`` (dataset_utils.py:67-74:
    ``soundfile.read`` scales PCM_16 by 1/32768, then ``.mean(1)``).  Half the bytes of the float32 path in HBM."""
    if not _is_int16(pcm):
        raise ValueError("pcm16_to_log_mel expects int16 samples")
    x = pcm if isinstance(pcm, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pcm))
    if x.dim() == 2:
        x = x[:, :, None]
    if x.dim() != 3:
        raise ValueError("pcm16_to_log_mel expects [B, samples] or [B, samples, channels]")
    B, n, C = x.shape
    if not 1 <= C <= 16:
        raise ValueError("1..16 interleaved channels are supported")
    x = (x.cuda() if not x.is_cuda else x).contiguous()
    norm = _norm_tensor(mean, std)
    out = torch.empty((B, num_frames(n), cfg.mel_bins), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _ext.check(_ext.load().sedb_logmel_pcm16(_ext.context(), _ptr(x), B, n, n, C, _ptr(norm), _ptr(out),
                                                 _ext.stream_ptr()))
    return out


def waveform_to_log_mel(wave, mean=None, std=None) -> torch.Tensor:
    """Fused log-mel of a batch of mono clips: ``[B, samples]`` (or ``[samples]``) -> ``[B, T, 64]`` float32 CUDA.
    int16 input is taken as 16-bit PCM (see :func:`pcm16_to_log_mel`)."""
    if _is_int16(wave):
        one = wave.ndim == 1
        out = pcm16_to_log_mel(wave[None] if one else wave, mean, std)
        return out[0] if one else out
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


# ------------------------------------------------------------------------------------------------------------------
# Callers of the hot path: offline feature extraction and the dataset transform, batched on
# the device.
def transform(x, mean, std, preprocessed_mode='logMel'):
    """Batched ``SpectogramDataset.transform`` (reference spectograms_dataset.py:104-110).

    ``logMel`` mode: ``(x - mean) / std`` with per-mel-bin statistics.  ``Complex`` mode: the complex STFT is normalised
    per FFT bin (complex mean, real std) and then converted with :func:`multichannel_complex_to_log_mel` -- one kernel
    call per batch instead of one NumPy call per sample inside DataLoader workers.
    """
    if preprocessed_mode not in ('logMel', 'Complex'):
        raise ValueError("Spectogram type should be either logmel or complex")
    is_np = not isinstance(x, torch.Tensor)
    xt = torch.from_numpy(np.ascontiguousarray(x)).cuda() if is_np else x
    m = torch.as_tensor(np.asarray(mean) if not isinstance(mean, torch.Tensor) else mean).to(xt.device)
    s = torch.as_tensor(np.asarray(std) if not isinstance(std, torch.Tensor) else std).to(xt.device)
    y = (xt - m) / s
    if preprocessed_mode == 'Complex':
        y = multichannel_complex_to_log_mel(y.to(torch.complex64))
    return y.cpu().numpy() if is_np else y


def preprocess_data(audio_path_and_labels, output_dir, output_mean_std_file, preprocess_mode='logMel',
                    read_audio=None, batch_files=16, pcm16=False):
    """Drop-in for the reference's ``preprocess_data`` (preprocess.py:60-81), batched through the fused kernel.

    Writes, per file, ``<audio_name>_<mode>_features_and_labels.pkl`` = ``{'features', 'start_times', 'end_times'}``
    with ``features`` a ``(C, T, 64)`` float32 log-mel (or ``(C, T, 16385)`` complex64 STFT in ``Complex`` mode) and the
    dataset-wide ``{'mean', 'std'}`` pickle (per feature bin over all channels and frames) -- the same on-disk format, so
    the reference's ``SpectogramDataset`` loads the result unchanged.  ``read_audio(path) -> (samples, channels)``
    defaults to the reference's ``read_multichannel_audio`` (needs ``soundfile``); files of equal length are processed
    ``batch_files`` at a time.  The debug plot of the reference (preprocess.py:83-86) is not produced.

    ``pcm16=True`` (logMel mode, ``audio_channels == 1``): 16-bit PCM WAV files are read as stored
    (``dataset_utils.read_wav_pcm16``) and go to the device as int16; the ``/ 32768`` scaling and the channel mean of
    ``read_multichannel_audio`` happen inside the log-mel kernel's loader (half the host->device bytes of float32 mono,
    an eighth of 4-channel float64).
    """
    import os
    import pickle

    if pcm16:
        from .. import dataset_utils
        if preprocess_mode != 'logMel' or cfg.audio_channels != 1:
            raise ValueError("pcm16=True needs preprocess_mode='logMel' and audio_channels == 1")

        def read_audio(path):                          # noqa: F811 - the int16 reader replaces the float one
            pcm, rate = dataset_utils.read_wav_pcm16(path)
            if rate != cfg.working_sample_rate:
                raise RuntimeError(f"{path}: sample rate {rate} != {cfg.working_sample_rate}; resample first")
            return pcm
    if read_audio is None:
        try:
            import soundfile  # noqa: F401
        except ImportError as e:                      # pragma: no cover - depends on the host
            raise RuntimeError("preprocess_data needs `soundfile` to decode audio (or pass read_audio=...)") from e
        read_audio = _read_multichannel_audio
    os.makedirs(output_dir, exist_ok=True)
    n_feat = cfg.mel_bins if preprocess_mode == 'logMel' else NUM_BINS
    s1 = torch.zeros(n_feat, dtype=torch.float64, device="cuda")     # per-bin running sums (complex handled apart)
    s2 = torch.zeros(n_feat, dtype=torch.float64, device="cuda")
    s1c = torch.zeros(n_feat, dtype=torch.complex128, device="cuda")
    count = 0

    def flush(group):
        nonlocal s1, s2, s1c, count
        waves = [g[0] for g in group]
        if pcm16:
            feats = pcm16_to_log_mel(torch.from_numpy(np.ascontiguousarray(np.stack(waves))).cuda())
            f64 = feats.to(torch.float64)
            s1 += f64.sum((0, 1))
            s2 += (f64 * f64).sum((0, 1))
            count += feats.shape[0] * feats.shape[1]
            feats = feats[:, None].cpu().numpy()                   # (files, 1, T, 64)
            for (wave, start_times, end_times, audio_name), feature in zip(group, feats):
                path = os.path.join(output_dir, audio_name + f"_{preprocess_mode}_features_and_labels.pkl")
                with open(path, 'wb') as fh:
                    pickle.dump({'features': feature, 'start_times': start_times, 'end_times': end_times}, fh)
            return
        C = waves[0].shape[1]
        stacked = torch.from_numpy(np.ascontiguousarray(np.stack([w.T for w in waves]), dtype=np.float32)).cuda()
        flat = stacked.reshape(len(waves) * C, -1)                   # one mono "clip" per (file, channel)
        if preprocess_mode == 'logMel':
            feats = waveform_to_log_mel(flat)
            f64 = feats.to(torch.float64)
            s1 += f64.sum((0, 1))
            s2 += (f64 * f64).sum((0, 1))
        else:
            feats = multichannel_stft(flat.t().contiguous())         # (files * C, T, 16385) complex64 on the device
            f = feats.to(torch.complex128)
            s1c += f.sum((0, 1))
            s2 += (f.real ** 2 + f.imag ** 2).sum((0, 1))
        count += feats.shape[0] * feats.shape[1]
        feats = feats.reshape(len(waves), C, feats.shape[-2], feats.shape[-1]).cpu().numpy()
        for (wave, start_times, end_times, audio_name), feature in zip(group, feats):
            path = os.path.join(output_dir, audio_name + f"_{preprocess_mode}_features_and_labels.pkl")
            with open(path, 'wb') as fh:
                pickle.dump({'features': feature, 'start_times': start_times, 'end_times': end_times}, fh)

    group = []
    for (audio_path, start_times, end_times, audio_name) in audio_path_and_labels:
        wave = np.asarray(read_audio(audio_path))
        if group and (wave.shape != group[0][0].shape or len(group) >= batch_files):
            flush(group)
            group = []
        group.append((wave, start_times, end_times, audio_name))
    if group:
        flush(group)

    if preprocess_mode == 'logMel':
        mean = s1 / count
        std = torch.sqrt(torch.clamp(s2 / count - mean * mean, min=0.0))
        mean_np, std_np = mean.cpu().numpy(), std.cpu().numpy()
    else:                                             # np.mean / np.std of a complex array: complex mean, real std
        mean_c = s1c / count
        std = torch.sqrt(torch.clamp(s2 / count - (mean_c.real ** 2 + mean_c.imag ** 2), min=0.0))
        mean_np, std_np = mean_c.cpu().numpy(), std.cpu().numpy()
    with open(output_mean_std_file, 'wb') as fh:
        pickle.dump({'mean': mean_np, 'std': std_np}, fh)
    return mean_np, std_np


def _read_multichannel_audio(audio_path, target_fs=None):
    """Reference ``read_multichannel_audio`` (dataset_utils.py:63-86) for files already at the working sample rate:
    decode with soundfile, downmix to ``audio_channels`` by averaging.  Resampling is out of scope."""
    import soundfile
    audio, sample_rate = soundfile.read(audio_path)
    if audio.ndim == 1:
        audio = audio.reshape(-1, 1)
    if audio.shape[1] < cfg.audio_channels:
        audio = np.repeat(audio.mean(1).reshape(-1, 1), cfg.audio_channels, axis=1)
    elif cfg.audio_channels == 1:
        audio = audio.mean(1).reshape(-1, 1)
    elif audio.shape[1] > cfg.audio_channels:
        audio = audio[:, :cfg.audio_channels]
    if sample_rate != cfg.working_sample_rate:
        raise RuntimeError(f"{audio_path}: sample rate {sample_rate} != {cfg.working_sample_rate}; resample first")
    return audio
