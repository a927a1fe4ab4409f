"""GPU parity of the 16-bit PCM entry points (sedb_logmel_pcm16 / sedb_logmel_host_pcm16 / sedb_sed_host_pcm16) against
the oracle: reference reader semantics (int16 / 32768, channel mean) followed by the reference log-mel.
Tolerance: log-mel within 1e-2 dB, probabilities within 1e-3 (north star)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import sed_b200
from sed_b200 import _ext
from sed_b200.dataset.spectogram import preprocess as P
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from oracle import audio_ref as A
from oracle import logmel_ref as R
import refmodels
import signals
from test_gpu_logmel import assert_parity

TOL_DB = 1e-2


def _pcm(n, channels, seed, gain=0.3):
    """A multichannel int16 recording: the parity signal with per-channel gains/delays plus independent noise."""
    rng = np.random.default_rng(seed)
    base = signals.hdr(n + 64, seed)
    chans = []
    for c in range(channels):
        x = gain * (0.6 + 0.2 * c) * base[c * 7:c * 7 + n] + 0.01 * rng.standard_normal(n)
        chans.append(np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16))
    return np.stack(chans, axis=1)


@pytest.mark.parametrize("channels", [1, 2, 4, 3, 6])
@pytest.mark.parametrize("n", [480000, 47521])
def test_pcm16_logmel_parity(channels, n):
    """1, 2 and 4 channels take the vector loads, 3 and 6 the general path; 47521 is an odd, unaligned length."""
    pcm = _pcm(n, channels, seed=channels)
    out = P.pcm16_to_log_mel(torch.from_numpy(pcm[None]).cuda()).cpu().numpy()[0]
    ref = R.waveform_to_log_mel(A.pcm16_to_mono(pcm))
    assert out.shape == ref.shape == (1 + n // 15840, 64)
    assert_parity(out, ref, TOL_DB)


def test_pcm16_batch_matches_float32_path_and_dispatch():
    """The int16 path equals the float32 path fed with the same mono mix (same kernel arithmetic after the loader)."""
    pcm = np.stack([_pcm(95040, 2, seed=s) for s in range(5)])                 # [5, n, 2]
    a = P.waveform_to_log_mel(torch.from_numpy(pcm).cuda()).cpu().numpy()       # dispatches on dtype
    mono = np.stack([A.pcm16_to_mono(p) for p in pcm]).astype(np.float32)
    b = P.waveform_to_log_mel(torch.from_numpy(mono).cuda()).cpu().numpy()
    assert a.shape == b.shape == (5, 7, 64)
    assert np.abs(a - b).max() < 2e-4
    one = P.waveform_to_log_mel(torch.from_numpy(pcm[0, :, 0].copy()).cuda())   # 1-D int16
    assert one.shape == (7, 64)


def test_pcm16_extreme_values_and_silence():
    n = 63360
    full = np.full((n, 1), -32768, dtype=np.int16)
    full[::2] = 32767
    out = P.pcm16_to_log_mel(torch.from_numpy(full[None]).cuda()).cpu().numpy()[0]
    assert_parity(out, R.waveform_to_log_mel(A.pcm16_to_mono(full)), TOL_DB)
    zero = np.zeros((1, n, 2), dtype=np.int16)
    out = P.pcm16_to_log_mel(torch.from_numpy(zero).cuda()).cpu().numpy()
    assert np.all(out == -100.0)                                               # 10 log10(amin = 1e-10)


def test_pcm16_rejects_bad_input():
    with pytest.raises(ValueError):
        P.pcm16_to_log_mel(torch.zeros(1, 40000, dtype=torch.float32))
    with pytest.raises(ValueError):
        P.pcm16_to_log_mel(torch.zeros(1, 40000, 17, dtype=torch.int16))
    with pytest.raises(RuntimeError):
        P.pcm16_to_log_mel(torch.zeros(1, 16000, dtype=torch.int16))           # shorter than the reflect padding


def test_pcm16_host_pipelines():
    lib = _ext.load()
    n, C, B = 190080, 2, 6
    pcm = np.stack([_pcm(n, C, seed=10 + s) for s in range(B)])
    host = torch.from_numpy(pcm).pin_memory()
    T = 1 + n // 15840
    out = torch.empty(B, T, 64).pin_memory()
    _ext.check(lib.sedb_logmel_host_pcm16(_ext.context(), ctypes.c_void_p(host.data_ptr()), B, n, n, C, None,
                                          ctypes.c_void_p(out.data_ptr())))
    dev = P.pcm16_to_log_mel(host.cuda()).cpu()
    assert torch.equal(out, dev)
    # end to end: PCM -> log-mel -> CNN probabilities, against the module fed with the oracle's log-mel
    model, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    model = model.cuda().eval()
    handle = model._handle(torch.device("cuda", torch.cuda.current_device()))
    frames = int(lib.sedb_cnn_out_frames(handle, T))
    probs = torch.empty(B, frames, 1).pin_memory()
    _ext.check(lib.sedb_sed_host_pcm16(_ext.context(), handle, ctypes.c_void_p(host.data_ptr()), B, n, n, C, None,
                                       ctypes.c_void_p(probs.data_ptr())))
    ref_lm = np.stack([R.waveform_to_log_mel(A.pcm16_to_mono(p)) for p in pcm]).astype(np.float32)
    with torch.no_grad():
        ref = model.logits(torch.from_numpy(ref_lm)[:, None].cuda()).cpu()
    assert (probs - ref).abs().max() < 1e-3
