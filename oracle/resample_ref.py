"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the resampling step in front of the hot path.

The reference resamples files whose rate differs from ``cfg.working_sample_rate`` with
``librosa.resample(x, orig_sr=fs, target_sr=target_fs)`` (dataset/dataset_utils.py:77-84), one channel at a time.
librosa is an un-vendored, UNPINNED dependency (README.md:11-13) and the default ``res_type`` of that call depends on
its version: ``kaiser_best`` (resampy's Kaiser-windowed sinc interpolation) before 0.10, ``soxr_hq`` from 0.10 on.
Neither library is installable here, so this row is PARITY UNPINNED against the reference.  What is restated is the
band-limited sinc interpolation  y[j] = sum_i x[i] h(i / orig - j / new)  with resampy's published ``kaiser_best``
design (64 zero crossings, roll-off 0.9475937167399596, Kaiser beta 14.769656459379492) in its exact polyphase form:
for reduced rates orig/new = L_o / L_n there are L_n distinct filters of 2 * width + L_o taps, evaluated in float64
(resampy itself interpolates a 512-per-zero-crossing table of the same filter linearly).  The restatement is pinned
against an independent implementation that IS installed, ``torchaudio.functional.resample(...,
resampling_method="sinc_interp_kaiser")`` with the same three parameters (tests/test_oracle_audio.py), and the
output length follows librosa: ceil(n * target_sr / orig_sr).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
import math

import numpy as np

LOWPASS_FILTER_WIDTH = 64
ROLLOFF = 0.9475937167399596
KAISER_BETA = 14.769656459379492


def reduced_rates(orig_sr, target_sr):
    g = math.gcd(int(orig_sr), int(target_sr))
    return int(orig_sr) // g, int(target_sr) // g


def polyphase_filters(orig_sr, target_sr):
    """(filters [L_n, 2 * width + L_o] float64, width): filters[p, k] weighs x[i * L_o + k - width] in y[i * L_n + p]."""
    lo, ln = reduced_rates(orig_sr, target_sr)
    base = min(lo, ln) * ROLLOFF
    width = int(math.ceil(LOWPASS_FILTER_WIDTH * lo / base))
    idx = np.arange(-width, width + lo, dtype=np.float64) / lo
    t = (-np.arange(ln, dtype=np.float64) / ln)[:, None] + idx[None, :]
    t = np.clip(t * base, -LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = np.i0(KAISER_BETA * np.sqrt(1.0 - (t / LOWPASS_FILTER_WIDTH) ** 2)) / np.i0(KAISER_BETA)
    return np.sinc(t) * window * (base / lo), width          # np.sinc(t) = sin(pi t) / (pi t)


def num_samples(n, orig_sr, target_sr):
    lo, ln = reduced_rates(orig_sr, target_sr)
    return (n * ln + lo - 1) // lo


def resample(x, orig_sr, target_sr):
    """1-D float64 resampling (dataset_utils.py:83 for one channel)."""
    x = np.asarray(x, dtype=np.float64)
    if orig_sr == target_sr:
        return x.copy()
    lo, ln = reduced_rates(orig_sr, target_sr)
    h, width = polyphase_filters(orig_sr, target_sr)
    n_out = num_samples(x.size, orig_sr, target_sr)
    n_blocks = (n_out + ln - 1) // ln
    xp = np.zeros(width + (n_blocks - 1) * lo + h.shape[1] + x.size, dtype=np.float64)
    xp[width:width + x.size] = x
    win = np.lib.stride_tricks.as_strided(xp, shape=(n_blocks, h.shape[1]), strides=(xp.strides[0] * lo, xp.strides[0]))
    return (win @ h.T).reshape(-1)[:n_out]
