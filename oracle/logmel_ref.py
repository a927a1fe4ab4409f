"""NumPy restatement of the reference's log-mel feature extraction (TEST INFRASTRUCTURE).

Follows ``dataset/spectogram/preprocess.py`` of the reference:

* ``MEL_FILTER_BANK_MATRIX``            preprocess.py:13-18  -> :func:`mel_filter_bank_matrix`
* ``multichannel_stft``                 preprocess.py:21-36  -> :func:`multichannel_stft`
* ``multichannel_complex_to_log_mel``   preprocess.py:39-45  -> :func:`multichannel_complex_to_log_mel`
* ``calculate_scalar_of_tensor``        preprocess.py:48-57  -> :func:`calculate_scalar_of_tensor`

The arithmetic itself lives in librosa (un-vendored, version unpinned in the
reference's README.md:11-13; the keyword call style matches librosa 0.8-0.11).
What is restated here is librosa's published algorithm for exactly the
call-site arguments:

* ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with the defaults
  ``htk=False, norm='slaney', dtype=float32``: Slaney mel scale (linear below
  1 kHz, log above), triangular filters from ``ramps/fdiff``, area
  normalisation ``2/(f[i+2]-f[i])``.
* ``librosa.core.stft(y, n_fft, hop_length, win_length, window, center=True,
  dtype=complex64, pad_mode='reflect')``: window centre-padded with zeros to
  ``n_fft``; signal reflect-padded by ``n_fft//2``; frame ``t`` starts at
  ``t*hop``; ``T = 1 + len(y)//hop``; rFFT in the input precision (float64
  for soundfile input); result cast to complex64.
* ``librosa.core.power_to_db(S, ref=1.0, amin=1e-10, top_db=None)``:
  ``10*log10(max(amin,S)) - 10*log10(max(amin,ref))``.

PARITY: unpinned by the reference (no tests, no fixtures upstream); pinned here
against torch.stft / torchaudio (tests/test_oracle_logmel.py).
"""
from __future__ import annotations

import numpy as np

# Constants of dataset/common_config.py:2-8 and dataset/spectogram/spectogram_configs.py:5-8
TIME_MARGIN = 0.33
SAMPLE_RATE = 48000
FRAME_SIZE = int(SAMPLE_RATE * TIME_MARGIN * 2)        # 31680
HOP_SIZE = FRAME_SIZE // 2                             # 15840
NFFT = 2 ** int(np.ceil(np.log2(FRAME_SIZE)))          # 32768
MEL_BINS = 64
MEL_MIN_FREQ = 20
MEL_MAX_FREQ = SAMPLE_RATE // 2


# ----------------------------------------------------------------------------------------------
# librosa.filters.mel restatement
# ----------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(freq):
    freq = np.asanyarray(freq, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (freq - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if freq.ndim:
        log_t = freq >= min_log_hz
        mels[log_t] = min_log_mel + np.log(freq[log_t] / min_log_hz) / logstep
    elif freq >= min_log_hz:
        mels = min_log_mel + np.log(freq / min_log_hz) / logstep
    return mels


def _mel_to_hz_slaney(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels, fmin, fmax):
    """Centre/edge frequencies (Hz) of the mel bands: librosa.mel_frequencies(htk=False)."""
    min_mel = _hz_to_mel_slaney(float(fmin))
    max_mel = _hz_to_mel_slaney(float(fmax))
    mels = np.linspace(min_mel, max_mel, n_mels)
    return _mel_to_hz_slaney(mels)


def librosa_filters_mel(sr, n_fft, n_mels, fmin, fmax):
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` -> (n_mels, 1+n_fft//2) float32."""
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]          # float32 *= float64 column, as librosa does
    return weights


def mel_filter_bank_matrix():
    """``MEL_FILTER_BANK_MATRIX`` of preprocess.py:13-18: (16385, 64) float32."""
    return librosa_filters_mel(SAMPLE_RATE, NFFT, MEL_BINS, MEL_MIN_FREQ, MEL_MAX_FREQ).T


# ----------------------------------------------------------------------------------------------
# librosa.core.stft restatement
# ----------------------------------------------------------------------------------------------
def padded_window(frame_size=FRAME_SIZE, n_fft=NFFT):
    """``np.hanning(win_length)`` centre-padded with zeros to ``n_fft`` (librosa util.pad_center)."""
    win = np.hanning(frame_size)
    lpad = (n_fft - frame_size) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + frame_size] = win
    return out


def librosa_stft(y, n_fft=NFFT, hop_length=HOP_SIZE, win_length=FRAME_SIZE, dtype=np.complex64):
    """``librosa.core.stft(y, n_fft, hop_length, win_length, window=np.hanning(win_length),
    center=True, dtype=complex64, pad_mode='reflect')`` -> (1+n_fft//2, T)."""
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("librosa_stft expects a mono signal")
    if not np.issubdtype(y.dtype, np.floating):
        y = y.astype(np.float64)
    if len(y) <= n_fft // 2:
        raise ValueError("reflect padding needs len(y) > n_fft//2")
    fft_window = padded_window(win_length, n_fft).astype(y.dtype)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop_length          # == 1 + len(y)//hop
    out = np.empty((1 + n_fft // 2, n_frames), dtype=dtype)
    for t in range(n_frames):                               # librosa blocks this loop by MAX_MEM_BLOCK
        seg = yp[t * hop_length: t * hop_length + n_fft]
        out[:, t] = np.fft.rfft(fft_window * seg)
    return out


def power_to_db(S, ref=1.0, amin=1e-10, top_db=None):
    """``librosa.core.power_to_db`` (preprocess.py:42-43 uses ref=1.0, amin=1e-10, top_db=None)."""
    S = np.asarray(S)
    magnitude = S
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------------------------
# dataset/spectogram/preprocess.py restatement
# ----------------------------------------------------------------------------------------------
_MEL = None


def _mel():
    global _MEL
    if _MEL is None:
        _MEL = mel_filter_bank_matrix()
    return _MEL


def multichannel_stft(multichannel_signal):
    """preprocess.py:21-36: (samples, C) -> (C, T, 16385) complex64."""
    multichannel_signal = np.asarray(multichannel_signal)
    (samples, channels_num) = multichannel_signal.shape
    features = []
    for c in range(channels_num):
        features.append(librosa_stft(multichannel_signal[:, c]).T)
    return np.array(features)


def multichannel_complex_to_log_mel(multichannel_complex_spectogram):
    """preprocess.py:39-45: complex (C,T,16385) or (T,16385) -> float32 log-mel (.., T, 64).

    The reference multiplies in float32 (``np.dot`` of float32 operands).  Accumulating in
    float64 and rounding once differs from any float32 summation order by <1e-5 relative
    (4e-5 dB), four orders below the 1e-2 dB parity tolerance.
    """
    power = np.abs(multichannel_complex_spectogram) ** 2          # float32 for complex64 input
    mel = np.dot(power.astype(np.float64), _mel().astype(np.float64)).astype(np.float32)
    return power_to_db(mel, ref=1.0, amin=1e-10, top_db=None).astype(np.float32)


def waveform_to_log_mel(wave):
    """(samples,) or (B, samples) float -> (T,64) / (B,T,64) float32: stft -> log-mel composed."""
    wave = np.asarray(wave, dtype=np.float64)
    if wave.ndim == 1:
        return multichannel_complex_to_log_mel(multichannel_stft(wave[:, None]))[0]
    return np.stack([multichannel_complex_to_log_mel(multichannel_stft(w[:, None]))[0] for w in wave])


def calculate_scalar_of_tensor(x):
    """preprocess.py:48-57."""
    if x.ndim == 2:
        axis = 0
    elif x.ndim == 3:
        axis = (0, 1)
    return np.mean(x, axis=axis), np.std(x, axis=axis)


def num_frames(n_samples, hop=HOP_SIZE):
    return 1 + n_samples // hop
