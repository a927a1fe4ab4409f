"""Pins oracle/logmel_ref.py (the CPU restatement of the reference's librosa calls at
dataset/spectogram/preprocess.py:13-45) against independent implementations and the committed vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_ref as R
import signals
import dft_model

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_config_constants():
    # dataset/common_config.py:2-8, dataset/spectogram/spectogram_configs.py:5-8
    assert (R.FRAME_SIZE, R.HOP_SIZE, R.NFFT, R.MEL_BINS) == (31680, 15840, 32768, 64)


def test_mel_matrix_structure():
    m = R.mel_filter_bank_matrix()
    assert m.shape == (16385, 64) and m.dtype == np.float32
    assert int((m != 0).sum()) == 31676
    nz0, nz63 = np.nonzero(m[:, 0])[0], np.nonzero(m[:, 63])[0]
    assert (nz0[0], nz0[-1], nz63[0], nz63[-1]) == (14, 98, 14403, 16384)
    assert abs(float(m.max()) - 0.015991725) < 1e-9
    assert np.all(m[:14] == 0)


def test_mel_matrix_vs_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(16385, 20.0, 24000.0, 64, 48000, norm="slaney", mel_scale="slaney").numpy()
    assert np.abs(fb - R.mel_filter_bank_matrix()).max() < 1e-7


@pytest.mark.parametrize("n", [31680, 100000])
def test_stft_vs_torch_stft(n):
    y = signals.hdr(n, 5)
    s = R.librosa_stft(y)
    st = torch.stft(torch.from_numpy(y), n_fft=32768, hop_length=15840, win_length=31680,
                    window=torch.from_numpy(np.hanning(31680)), center=True, pad_mode="reflect",
                    return_complex=True).numpy()
    assert s.shape == st.shape == (16385, 1 + n // 15840)
    assert s.dtype == np.complex64
    assert np.abs(s - st).max() / np.abs(st).max() < 1e-6


@pytest.mark.parametrize("n,t", [(31680, 3), (480000, 31), (2880000, 182), (2880001, 182)])
def test_num_frames(n, t):
    assert R.num_frames(n) == t


def test_shapes_and_2d_input():
    y = signals.white(50000, 1)
    spec = R.multichannel_stft(y[:, None])
    assert spec.shape == (1, 4, 16385) and spec.dtype == np.complex64
    lm3 = R.multichannel_complex_to_log_mel(spec)
    lm2 = R.multichannel_complex_to_log_mel(spec[0])          # Classical_methods/train_svm_detector.py:68
    assert lm3.shape == (1, 4, 64) and lm2.shape == (4, 64) and lm3.dtype == np.float32
    assert np.array_equal(lm3[0], lm2)


def test_power_to_db_floor():
    assert R.power_to_db(np.array([0.0, 1e-12, 1.0, 100.0])).tolist() == [-100.0, -100.0, 0.0, 20.0]


def test_tone_peak_location():
    lm = R.waveform_to_log_mel(signals.tone(48000, 1000.0))
    mel_f = R.mel_frequencies(66, 20, 24000)
    expect = int(np.argmin(np.abs(mel_f[1:-1] - 1000.0)))
    assert int(np.argmax(lm[1])) == expect


def test_golden_vectors():
    gold = np.load(os.path.join(GOLD, "logmel_oracle.npz"))
    for name, fn in signals.ALL.items():
        assert np.abs(R.waveform_to_log_mel(fn(100000, 3)) - gold[f"{name}_100000"]).max() < 1e-4
    assert np.abs(R.waveform_to_log_mel(signals.tone(48000)) - gold["tone1k_48000"]).max() < 1e-4
    assert np.abs(R.waveform_to_log_mel(signals.impulse(31680, 0)) - gold["impulse0_31680"]).max() < 1e-4


def test_calculate_scalar_of_tensor():
    x = np.random.default_rng(0).standard_normal((2, 5, 64)).astype(np.float32)
    mean, std = R.calculate_scalar_of_tensor(x)
    assert mean.shape == std.shape == (64,)
    assert np.allclose(mean, x.reshape(-1, 64).mean(0), atol=1e-6)


# ---- the factored-DFT dataflow the CUDA kernel implements (tests/dft_model.py) --------------------------
@pytest.mark.parametrize("name", ["white", "hdr", "silence"])
def test_factored_dft_matches_rfft(name):
    g = R.padded_window() * signals.ALL[name](32768, 2)
    p, x = dft_model.factored_power_spectrum(g, None)
    xr = np.fft.rfft(g)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-6
    assert np.abs(p - np.abs(xr) ** 2).max() / (np.abs(xr) ** 2).max() < 1e-6


@pytest.mark.parametrize("rnd", [dft_model.bf16_round, dft_model.fp16_round])
def test_split_operand_accuracy(rnd):
    """hi*hi + lo*hi + hi*lo split products keep the log-mel far inside the 1e-2 dB tolerance."""
    mel = R.mel_filter_bank_matrix().astype(np.float64)
    g = R.padded_window() * signals.hdr(32768, 2)
    p, _ = dft_model.factored_power_spectrum(g, rnd)
    ref = np.abs(np.fft.rfft(g)) ** 2
    db = 10 * np.log10(np.maximum(1e-10, p @ mel))
    dbr = 10 * np.log10(np.maximum(1e-10, ref @ mel))
    assert np.abs(db - dbr).max() < 1e-3


# ---- numerics of the v2 dataflow (fold + radix-2) as the kernel runs it: split operands, per-frame block scale ----
def _block_scale(g):
    """2^e with e = 5 - floor(log2 max|g|) rounded down to even (csrc/logmel.cuh)."""
    mx = np.abs(g).max()
    if mx == 0:
        return 1.0
    e = (5 - int(np.floor(np.log2(mx)))) & ~1
    return float(2.0 ** max(-56, min(60, e)))


def _mel_db(p):
    return 10 * np.log10(np.maximum(1e-10, p @ R.mel_filter_bank_matrix().astype(np.float64)))


def test_v2_dataflow_matches_rfft_exactly_in_float64():
    g = R.padded_window() * signals.hdr(32768, 5)
    p, x = dft_model.factored_power_spectrum_v2(g, None)
    xr = np.fft.rfft(g)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-6


@pytest.mark.parametrize("rnd,window_db", [(dft_model.fp16_round, 100.0), (dft_model.bf16_round, 75.0)])
def test_dynamic_range_contract_on_a_pure_tone(rnd, window_db):
    """A pure tone spans > 150 dB per frame in float64.  The split-operand DFT keeps the 1e-2 dB bound for every mel
    bin within `window_db` of the loudest one and keeps the others below that window."""
    t = np.arange(32768) / 48000.0
    g = R.padded_window() * (0.5 * np.sin(2 * np.pi * 1000.0 * t))
    scale = _block_scale(g) if rnd is dft_model.fp16_round else 1.0
    p, _ = dft_model.factored_power_spectrum_v2(g, rnd, scale)
    db, ref = _mel_db(p), _mel_db(np.abs(np.fft.rfft(g)) ** 2)
    live = ref > ref.max() - window_db
    assert live.sum() >= 2 and (~live).sum() >= 8            # the tone really exceeds the window
    assert np.abs(db - ref)[live].max() < 1e-2
    assert np.all(db[~live] < ref.max() - window_db + 3.0)


def test_block_scale_is_what_keeps_quiet_frames_accurate_in_fp16():
    """1e-4 amplitude: without the power-of-two block scale the fp16 low halves sink into subnormals."""
    g = R.padded_window() * (1e-4 * signals.white(32768, 9) / 0.1)
    ref = _mel_db(np.abs(np.fft.rfft(g)) ** 2)
    scaled, _ = dft_model.factored_power_spectrum_v2(g, dft_model.fp16_round, _block_scale(g))
    plain, _ = dft_model.factored_power_spectrum_v2(g, dft_model.fp16_round, 1.0)
    err_scaled, err_plain = np.abs(_mel_db(scaled) - ref).max(), np.abs(_mel_db(plain) - ref).max()
    assert err_scaled < 1e-3
    assert err_plain > 5 * err_scaled
    assert 16.0 <= np.abs(g).max() * _block_scale(g) < 64.0
