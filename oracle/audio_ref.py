"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the audio hand-over in front of the hot path.

Follows dataset/dataset_utils.py:63-76 of the reference (``read_multichannel_audio`` without the resampling branch)
for 16-bit PCM input: ``soundfile.read`` returns ``int16 / 32768`` as float64, then the channel policy for the
configured ``audio_channels``.  Parity unpinned: the reference ships no fixtures for it; the arithmetic is exact in
float64, so the restatement is pinned by construction (tests/test_oracle_audio.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
import numpy as np


def pcm16_to_float64(pcm):
    """soundfile.read(dtype='float64') of a PCM_16 file: samples / 2**15 (dataset_utils.py:67)."""
    pcm = np.asarray(pcm)
    if pcm.dtype != np.int16:
        raise ValueError("expected int16 PCM")
    return pcm.astype(np.float64) / 32768.0


def channel_policy(audio, audio_channels=1):
    """dataset_utils.py:68-76."""
    a = np.asarray(audio, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.shape[1] < audio_channels:
        a = np.repeat(a.mean(1).reshape(-1, 1), audio_channels, axis=1)
    elif audio_channels == 1:
        a = a.mean(1).reshape(-1, 1)
    elif a.shape[1] > audio_channels:
        a = a[:, :audio_channels]
    return a


def pcm16_to_mono(pcm):
    """(samples, channels) int16 -> (samples,) float64 mono as the reference feeds it to multichannel_stft."""
    return channel_policy(pcm16_to_float64(pcm), 1)[:, 0]
