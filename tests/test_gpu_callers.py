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
