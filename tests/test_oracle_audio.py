"""CPU: the audio hand-over in front of the hot path (reference dataset_utils.py:63-76) -- oracle restatement, the
package's reader mirror and the 16-bit PCM WAV decoder (standard library only, no GPU)."""
import os
import sys
import wave

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sed_b200  # noqa: E402,F401
from sed_b200.dataset import dataset_utils as DU  # noqa: E402
from sed_b200.dataset import common_config as cfg  # noqa: E402
from oracle import audio_ref as A  # noqa: E402


def _write_wav(path, pcm, rate=48000):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


def test_pcm16_scaling_is_soundfiles():
    pcm = np.array([-32768, -1, 0, 1, 32767], dtype=np.int16)
    np.testing.assert_array_equal(A.pcm16_to_float64(pcm), np.array([-1.0, -1 / 32768, 0.0, 1 / 32768, 32767 / 32768]))
    with pytest.raises(ValueError):
        A.pcm16_to_float64(pcm.astype(np.int32))


@pytest.mark.parametrize("channels", [1, 2, 3, 4])
def test_channel_policy_mono_config(channels):
    rng = np.random.default_rng(channels)
    pcm = rng.integers(-32768, 32768, size=(1000, channels), dtype=np.int16)
    mono = A.pcm16_to_mono(pcm)
    ref = (pcm.astype(np.float64) / 32768.0).mean(1)
    np.testing.assert_array_equal(mono, ref)
    np.testing.assert_array_equal(DU.apply_channel_policy(pcm.astype(np.float64) / 32768.0)[:, 0], ref)


def test_channel_policy_other_branches(monkeypatch):
    x = np.arange(12, dtype=np.float64).reshape(4, 3)
    assert A.channel_policy(x, 2).tolist() == x[:, :2].tolist()                 # more channels than configured
    rep = A.channel_policy(x[:, :1], 2)                                        # fewer: the mean, repeated
    assert rep.shape == (4, 2) and np.array_equal(rep[:, 0], x[:, 0]) and np.array_equal(rep[:, 1], x[:, 0])
    assert A.channel_policy(x[:, 0], 1).shape == (4, 1)                        # 1-D input
    monkeypatch.setattr(cfg, "audio_channels", 2)
    np.testing.assert_array_equal(DU.apply_channel_policy(x), x[:, :2])
    np.testing.assert_array_equal(DU.apply_channel_policy(x[:, :1]), rep)


@pytest.mark.parametrize("channels", [1, 4])
def test_wav_reader_round_trip(tmp_path, channels):
    rng = np.random.default_rng(7)
    pcm = rng.integers(-32768, 32768, size=(4801, channels), dtype=np.int16)
    path = tmp_path / "a.wav"
    _write_wav(path, pcm)
    got, rate = DU.read_wav_pcm16(str(path))
    assert rate == 48000 and got.dtype == np.int16
    np.testing.assert_array_equal(got, pcm)
    audio = DU.read_multichannel_audio(str(path), target_fs=cfg.working_sample_rate)
    assert audio.shape == (4801, 1) and audio.dtype == np.float64
    np.testing.assert_allclose(audio[:, 0], A.pcm16_to_mono(pcm), rtol=0, atol=1e-15)
    with pytest.raises(RuntimeError):
        DU.read_multichannel_audio(str(path), target_fs=44100)


def test_wav_reader_rejects_other_sample_widths(tmp_path):
    path = tmp_path / "b.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(1)
        w.setframerate(48000)
        w.writeframes(bytes(100))
    with pytest.raises(ValueError):
        DU.read_wav_pcm16(str(path))
