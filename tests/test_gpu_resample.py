"""GPU parity of the sample-rate converter (sedb_resample_f32 through the ctypes shim) against the CPU oracle
(oracle/resample_ref.py: Kaiser-windowed sinc, resampy's kaiser_best design, pinned against torchaudio on the CPU side).
Tolerance: 2e-5 absolute on samples in [-1, 1] (float32 accumulation of up to ~410 taps against float64)."""
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import sed_b200  # noqa: F401
from sed_b200.dataset import dataset_utils as DU
from sed_b200.dataset.spectogram import preprocess as P
from oracle import resample_ref as R
from oracle import logmel_ref
import signals

TOL = 2e-5


@pytest.mark.parametrize("orig,new", [(44100, 48000), (32000, 48000), (16000, 48000), (96000, 48000), (22050, 48000),
                                      (8000, 48000), (48000, 16000), (48000, 44100)])
@pytest.mark.parametrize("n", [1, 147, 20011])
def test_resample_matches_oracle(orig, new, n):
    x = np.random.default_rng(n + orig).uniform(-1, 1, n)
    y = DU.resample(x, orig, new)
    ref = R.resample(x, orig, new)
    assert y.dtype == np.float64 and y.shape == ref.shape
    assert np.abs(y - ref).max() < TOL


def test_resample_batch_strided_cuda_tensor():
    """[B, n] CUDA input with a row stride (a view of a wider buffer), several clips, result stays on the device."""
    rng = np.random.default_rng(3)
    buf = torch.from_numpy(rng.uniform(-1, 1, (5, 30000)).astype(np.float32)).cuda()
    x = buf[:, 17:17 + 26460]                                   # 0.6 s at 44.1 kHz, unaligned start
    y = DU.resample(x, 44100, 48000)
    assert y.is_cuda and y.dtype == torch.float32 and y.shape == (5, 28800)
    for b in range(5):
        assert np.abs(y[b].cpu().numpy() - R.resample(x[b].cpu().numpy(), 44100, 48000)).max() < TOL


def test_same_rate_is_a_copy_and_bad_arguments():
    x = np.random.default_rng(0).standard_normal(1000)
    np.testing.assert_array_equal(DU.resample(x.astype(np.float32), 48000, 48000), x.astype(np.float32))
    with pytest.raises(TypeError):
        DU.resample(np.zeros(10, dtype=np.int16), 44100, 48000)
    with pytest.raises(ValueError):
        DU.resample(x, 0, 48000)
    with pytest.raises(RuntimeError):
        DU.resample(x, 44101, 48000)                            # reduces to 44101 : 48000, more than 4096 phases


def test_read_multichannel_audio_resamples_like_the_reference_call(tmp_path):
    """dataset_utils.py:63-84 end to end: a 2-channel 44.1 kHz PCM_16 file -> channel mean -> 48 kHz, and the log-mel of
    the result equals the oracle's log-mel of the oracle's resampling."""
    rng = np.random.default_rng(9)
    pcm = (np.clip(signals.hdr(44100 * 2, 5), -1, 1)[:, None] * np.array([20000, 12000])[None, :]
           + rng.integers(-50, 50, (44100 * 2, 2))).astype(np.int16)
    path = os.path.join(tmp_path, "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.tobytes())
    audio = DU.read_multichannel_audio(path, target_fs=48000)
    mono = pcm.astype(np.float64).mean(1) / 32768.0
    ref = R.resample(mono, 44100, 48000)
    assert audio.shape == (96000, 1) and audio.dtype == np.float64
    assert np.abs(audio[:, 0] - ref).max() < TOL
    lm = P.waveform_to_log_mel(torch.from_numpy(audio[:, 0]).float().cuda()).cpu().numpy()
    assert np.abs(lm - logmel_ref.waveform_to_log_mel(ref)).max() < 1e-2
