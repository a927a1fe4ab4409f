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


# ---- resampling (reference dataset_utils.py:77-84) -----------------------------------------------------------------
RATE_PAIRS = [(44100, 48000), (32000, 48000), (16000, 48000), (96000, 48000), (22050, 48000), (8000, 48000),
              (48000, 16000)]


@pytest.mark.parametrize("orig,new", RATE_PAIRS)
def test_resample_oracle_is_pinned_by_torchaudio(orig, new):
    """The oracle restates a Kaiser-windowed sinc interpolation with resampy's kaiser_best parameters; the independent
    implementation installed here is torchaudio's sinc_interp_kaiser with the same three parameters."""
    torch = pytest.importorskip("torch")
    F = pytest.importorskip("torchaudio.functional")
    from oracle import resample_ref as R
    x = np.random.default_rng(orig + new).standard_normal(20011)
    y = R.resample(x, orig, new)
    t = F.resample(torch.from_numpy(x), orig, new, lowpass_filter_width=R.LOWPASS_FILTER_WIDTH, rolloff=R.ROLLOFF,
                   resampling_method="sinc_interp_kaiser", beta=R.KAISER_BETA).numpy()
    assert y.shape == t.shape == (-(-20011 * new // orig),)            # librosa: ceil(n * target_sr / orig_sr)
    assert np.abs(y - t).max() < 5e-7


def test_resample_oracle_properties():
    from oracle import resample_ref as R
    x = np.random.default_rng(1).standard_normal(1000)
    np.testing.assert_array_equal(R.resample(x, 48000, 48000), x)
    # a tone well below both Nyquist rates is reproduced at the new rate (away from the clip's ends)
    n, f = 44100, 1000.0
    y = R.resample(np.sin(2 * np.pi * f * np.arange(n) / 44100), 44100, 48000)
    ref = np.sin(2 * np.pi * f * np.arange(y.size) / 48000)
    assert y.size == 48000 and np.abs(y - ref)[200:-200].max() < 1e-4
    # unit DC gain of every phase
    h, _ = R.polyphase_filters(44100, 48000)
    assert np.abs(h.sum(1) - 1.0).max() < 1e-4


@pytest.mark.parametrize("orig,new", RATE_PAIRS)
def test_library_filter_table_matches_oracle(orig, new):
    """host_tables.h: make_resample_filters (float64 on the host, stored [tap][phase] float32) == the oracle's filters."""
    import ctypes
    from sed_b200 import _ext
    from oracle import resample_ref as R
    lib = _ext.load()
    w, t, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _ext.check(lib.sedb_resample_filters(orig, new, None, ctypes.byref(w), ctypes.byref(t), ctypes.byref(p)))
    h, width = R.polyphase_filters(orig, new)
    assert (p.value, t.value, w.value) == (h.shape[0], h.shape[1], width)
    tab = np.empty((t.value, p.value), dtype=np.float32)
    _ext.check(lib.sedb_resample_filters(orig, new, ctypes.c_void_p(tab.ctypes.data), None, None, None))
    assert np.abs(tab.T.astype(np.float64) - h).max() < 1e-7
    assert lib.sedb_resample_num_samples(20011, orig, new) == R.num_samples(20011, orig, new)
