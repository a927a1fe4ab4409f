"""Audio reading as the hot path's callers see it (reference: dataset/dataset_utils.py:63-86).

``read_multichannel_audio`` keeps the reference's name, arguments and channel handling.  Decoding uses ``soundfile``
when it is installed and falls back to the standard library's ``wave`` module for 16-bit PCM WAV files (the format of
both datasets the reference trains on).  ``read_wav_pcm16`` returns the samples as stored -- int16, interleaved --
for :func:`sed_b200.dataset.spectogram.preprocess.pcm16_to_log_mel`, which fuses the ``/ 32768`` scaling and the channel
mean into the log-mel kernel's loader.  Files at another rate are converted on the GPU by :func:`resample`
(``sedb_resample_f32``: Kaiser-windowed sinc, resampy's ``kaiser_best`` design), the counterpart of the reference's
``librosa.resample`` call (dataset_utils.py:77-84; parity unpinned: the reference pins no librosa version and newer
versions default to ``soxr_hq``).
"""
import wave

import numpy as np

from . import common_config as cfg


def read_wav_pcm16(audio_path):
    """``(samples, channels)`` int16 array and the sample rate of a 16-bit PCM WAV file."""
    with wave.open(audio_path, "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError(f"{audio_path}: only uncompressed 16-bit PCM WAV is supported "
                             f"(sample width {w.getsampwidth()}, compression {w.getcomptype()})")
        channels, rate, frames = w.getnchannels(), w.getframerate(), w.getnframes()
        data = np.frombuffer(w.readframes(frames), dtype="<i2")
    return data.reshape(-1, channels), rate


def apply_channel_policy(multichannel_audio):
    """dataset_utils.py:68-76: mono files become (samples, 1); fewer channels than configured -> the mean repeated;
    ``audio_channels == 1`` -> the channel mean; more channels than configured -> the first ``audio_channels``."""
    a = np.asarray(multichannel_audio)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.shape[1] < cfg.audio_channels:
        a = np.repeat(a.mean(1).reshape(-1, 1), cfg.audio_channels, axis=1)
    elif cfg.audio_channels == 1:
        a = a.mean(1).reshape(-1, 1)
    elif a.shape[1] > cfg.audio_channels:
        a = a[:, :cfg.audio_channels]
    return a


def resample(audio, orig_sr, target_sr):
    """``librosa.resample(audio, orig_sr=..., target_sr=...)`` for the last axis of ``audio`` on the GPU: NumPy in ->
    NumPy out (the input's floating dtype), CUDA tensor in -> float32 CUDA tensor out.  The arithmetic is float32."""
    import torch
    from .. import _ext
    is_np = not isinstance(audio, torch.Tensor)
    x = torch.as_tensor(np.ascontiguousarray(audio) if is_np else audio)
    if not x.is_floating_point():
        raise TypeError("resample expects floating-point samples")
    if int(orig_sr) != orig_sr or int(target_sr) != target_sr or orig_sr <= 0 or target_sr <= 0:
        raise ValueError("sample rates must be positive integers")
    out_dtype = x.dtype
    lead = x.shape[:-1]
    w = x.reshape(-1, x.shape[-1]).to(device="cuda" if not x.is_cuda else x.device, dtype=torch.float32).contiguous()
    lib = _ext.load()
    n_out = int(lib.sedb_resample_num_samples(w.shape[1], int(orig_sr), int(target_sr)))
    out = torch.empty((w.shape[0], n_out), dtype=torch.float32, device=w.device)
    if w.numel():
        with torch.cuda.device(w.device):
            _ext.check(lib.sedb_resample_f32(_ext.context(), w.data_ptr(), w.shape[0], w.shape[1], w.stride(0),
                                             int(orig_sr), int(target_sr), out.data_ptr(), out.stride(0),
                                             _ext.stream_ptr()))
    out = out.reshape(*lead, n_out)
    return out.cpu().numpy().astype(_np_dtype(out_dtype)) if is_np else out


def _np_dtype(torch_dtype):
    import torch
    return {torch.float64: np.float64, torch.float32: np.float32, torch.float16: np.float16}.get(torch_dtype, np.float32)


def read_multichannel_audio(audio_path, target_fs=None):
    """Reference signature and result: float64 ``(samples, audio_channels)`` at ``target_fs``."""
    try:
        import soundfile
        audio, sample_rate = soundfile.read(audio_path)
    except ImportError:
        pcm, sample_rate = read_wav_pcm16(audio_path)
        audio = pcm.astype(np.float64) / 32768.0               # soundfile's scaling of PCM_16
    audio = apply_channel_policy(audio)
    if target_fs is not None and sample_rate != target_fs:
        # dataset_utils.py:77-84: every channel resampled on its own
        audio = np.ascontiguousarray(resample(np.ascontiguousarray(audio.T), sample_rate, target_fs).T)
    return audio
