"""Audio reading as the hot path's callers see it (reference: dataset/dataset_utils.py:63-86).

``read_multichannel_audio`` keeps the reference's name, arguments and channel handling.  Decoding uses ``soundfile``
when it is installed and falls back to the standard library's ``wave`` module for 16-bit PCM WAV files (the format of
both datasets the reference trains on).  ``read_wav_pcm16`` returns the samples as stored -- int16, interleaved --
for :func:`sed_b200.dataset.spectogram.preprocess.pcm16_to_log_mel`, which fuses the ``/ 32768`` scaling and the channel
mean into the log-mel kernel's loader.  Resampling (``librosa.resample``, dataset_utils.py:77-84) is not provided:
files must already be at ``target_fs``.
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
        raise RuntimeError(f"{audio_path}: sample rate {sample_rate} != {target_fs}; resample the file first "
                           "(librosa.resample is outside this package)")
    return audio
