"""GPU tests of the callers either side of the hot path: offline feature extraction with
dataset-wide statistics (preprocess.py:60-81) and the batched dataset transform (spectograms_dataset.py:104-110)."""
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import sed_b200
from sed_b200.dataset.spectogram import preprocess as P
from oracle import logmel_ref as R
import signals


def test_preprocess_data_pickles_and_statistics(tmp_path):
    audio = {f"clip{i}": (signals.hdr(100000 + (7 if i == 3 else 0), 50 + i) * (0.3 + 0.2 * i))[:, None] for i in range(5)}
    items = [(name, [1.0 * i], [1.5 * i], name) for i, name in enumerate(audio)]
    out_dir, stats = str(tmp_path / "feat"), str(tmp_path / "mean_std.pkl")
    mean, std = P.preprocess_data(items, out_dir, stats, preprocess_mode="logMel", read_audio=lambda p: audio[p],
                                  batch_files=2)
    feats = []
    for i, name in enumerate(audio):
        d = pickle.load(open(os.path.join(out_dir, name + "_logMel_features_and_labels.pkl"), "rb"))
        assert set(d) == {"features", "start_times", "end_times"} and d["start_times"] == [1.0 * i]
        ref = R.multichannel_complex_to_log_mel(R.multichannel_stft(audio[name]))
        assert d["features"].shape == ref.shape and d["features"].dtype == np.float32
        assert np.abs(d["features"] - ref).max() < 1e-2
        feats.append(ref)
    ref_mean, ref_std = R.calculate_scalar_of_tensor(np.concatenate(feats, axis=1))
    d = pickle.load(open(stats, "rb"))
    assert d["mean"].shape == d["std"].shape == (64,)
    assert np.abs(d["mean"] - ref_mean).max() < 1e-2 and np.abs(d["std"] - ref_std).max() < 1e-2
    assert np.allclose(mean, d["mean"]) and np.allclose(std, d["std"])


def test_transform_logmel_and_complex_modes():
    y = signals.hdr(100000, 9)
    spec = R.multichannel_stft(y[:, None])                        # (1, 7, 16385) complex64
    lm = R.multichannel_complex_to_log_mel(spec)
    mean, std = lm.mean((0, 1)), lm.std((0, 1))
    out = P.transform(lm, mean, std, "logMel")
    assert np.abs(out - (lm - mean) / std).max() < 1e-5
    cmean, cstd = spec.mean((0, 1)), spec.std((0, 1)) + 1e-3
    ref = R.multichannel_complex_to_log_mel(((spec - cmean) / cstd).astype(np.complex64))
    out_c = P.transform(spec, cmean, cstd, "Complex")
    assert out_c.shape == ref.shape and np.abs(out_c - ref).max() < 1e-2
    with pytest.raises(ValueError):
        P.transform(lm, mean, std, "other")


def test_preprocess_data_from_pcm16_wav_files(tmp_path):
    """The f-3 hand-over: 16-bit PCM WAV files (4-channel, TAU-FOA-like) -> int16 batches -> channel mean and scaling
    inside the kernel loader; pickles and statistics equal the reference reader + reference log-mel."""
    import wave
    from oracle import audio_ref as A
    rng = np.random.default_rng(3)
    items, pcms = [], {}
    for i in range(3):
        n = 110880
        base = signals.hdr(n, 70 + i)
        pcm = np.stack([np.clip(np.round((0.2 + 0.1 * c) * base * 32768 + 20 * rng.standard_normal(n)), -32768, 32767)
                        for c in range(4)], axis=1).astype(np.int16)
        path = str(tmp_path / f"foa{i}.wav")
        with wave.open(path, "wb") as w:
            w.setnchannels(4); w.setsampwidth(2); w.setframerate(48000); w.writeframes(pcm.tobytes())
        items.append((path, [0.5], [1.0], f"foa{i}"))
        pcms[f"foa{i}"] = pcm
    out_dir, stats = str(tmp_path / "feat"), str(tmp_path / "mean_std.pkl")
    mean, std = P.preprocess_data(items, out_dir, stats, preprocess_mode="logMel", pcm16=True, batch_files=2)
    refs = []
    for name, pcm in pcms.items():
        d = pickle.load(open(os.path.join(out_dir, name + "_logMel_features_and_labels.pkl"), "rb"))
        ref = R.waveform_to_log_mel(A.pcm16_to_mono(pcm))[None]
        assert d["features"].shape == ref.shape == (1, 8, 64) and d["features"].dtype == np.float32
        assert np.abs(d["features"] - ref).max() < 1e-2
        refs.append(ref)
    ref_mean, ref_std = R.calculate_scalar_of_tensor(np.concatenate(refs, axis=1))
    assert np.abs(mean - ref_mean).max() < 1e-2 and np.abs(std - ref_std).max() < 1e-2
    with pytest.raises(ValueError):
        P.preprocess_data(items, out_dir, stats, preprocess_mode="Complex", pcm16=True)
