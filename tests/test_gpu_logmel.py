"""GPU parity of the log-mel path (through the C ABI via the ctypes shim) against the CPU oracle.
Tolerance (north star): log-mel within 1e-2 dB max-abs of the reference CPU path."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import sed_b200
from sed_b200 import _ext
from sed_b200.dataset.spectogram import preprocess as P
from oracle import logmel_ref as R
import signals

TOL_DB = 1e-2
# The reference DFT runs in float64; the tensor-core DFT carries ~2^-17 (bf16 split) / ~2^-22 (fp16 split) relative
# error against the frame's total energy.  The 1e-2 dB bound therefore holds for every mel bin within DYN_RANGE_DB of
# the loudest bin of the same frame (DESIGN.md "dynamic-range contract"); quieter bins must stay below that window.
DYN_RANGE_DB = 100.0 if _ext.load().sedb_split_is_fp16() else 75.0
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def assert_parity(out, ref, tol=TOL_DB):
    assert out.shape == ref.shape
    top = ref.max(axis=-1, keepdims=True)
    live = ref > top - DYN_RANGE_DB
    assert np.abs(out - ref)[live].max() < tol
    if (~live).any():
        assert np.all((out < top - DYN_RANGE_DB + 3.0)[~live])


def gpu_logmel(y, **kw):
    return P.waveform_to_log_mel(torch.from_numpy(np.asarray(y, dtype=np.float32)).cuda(), **kw).cpu().numpy()


@pytest.mark.parametrize("N,K,am,bm,pad,neg", [(128, 16, 0, 0, 0, 0), (128, 64, 0, 1, 0, 0), (256, 64, 0, 0, 0, 0),
                                               (128, 32, 0, 1, 16, 0), (64, 32, 1, 0, 0, 0), (32, 48, 0, 0, 0, 1),
                                               (128, 64, 2, 0, 0, 0), (128, 32, 2, 1, 0, 0)])
def test_umma_descriptor_conventions(N, K, am, bm, pad, neg):
    """Pins the tcgen05 shared-memory descriptor conventions the kernels rely on (see csrc/umma.cuh); am == 2 feeds
    the A operand from tensor memory (tcgen05.st + the [a_tmem] MMA form)."""
    g = torch.Generator().manual_seed(N + K)
    a = torch.randn(128, K, generator=g).cuda()
    b = torch.randn(K, N, generator=g).cuda()
    d = torch.zeros(128, N, device="cuda")
    _ext.check(_ext.load().sedb_debug_umma_probe(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                                 ctypes.c_void_p(d.data_ptr()), N, K, am, bm, pad, neg, 0, None))
    torch.cuda.synchronize()
    ref = a.bfloat16().float() @ b.bfloat16().float()
    if neg:
        ref = -ref
    assert (d - ref).abs().max() / ref.abs().max() < 1e-5


@pytest.mark.parametrize("name", ["white", "hdr", "silence"])
@pytest.mark.parametrize("n", [31680, 480000])
def test_logmel_parity_signal_classes(name, n):
    y = signals.ALL[name](n, 0)
    out, ref = gpu_logmel(y), R.waveform_to_log_mel(y)
    assert out.shape == ref.shape == (1 + n // 15840, 64) and out.dtype == np.float32
    assert np.abs(out - ref).max() < TOL_DB


@pytest.mark.parametrize("n", [2880000, 2880001, 16385, 47519, 47520])
def test_logmel_parity_lengths(n):
    """TAU-shaped 60 s clip, an unaligned length (scalar-load path), the shortest legal clip, frame-count edges."""
    y = signals.hdr(n, 1)
    out, ref = gpu_logmel(y), R.waveform_to_log_mel(y)
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() < TOL_DB


def test_logmel_golden_vectors():
    gold = np.load(os.path.join(GOLD, "logmel_oracle.npz"))
    for name, fn in signals.ALL.items():
        assert np.abs(gpu_logmel(fn(100000, 3)) - gold[f"{name}_100000"]).max() < TOL_DB
    # a pure tone spans > 150 dB per frame in float64: parity inside the dynamic-range window
    assert_parity(gpu_logmel(signals.tone(48000)), gold["tone1k_48000"])
    # impulse at sample 0 exercises the reflect padding (flat spectrum in frame 0, silence afterwards)
    out = gpu_logmel(signals.impulse(31680, 0))
    g = gold["impulse0_31680"]
    assert np.abs(out[0] - g[0]).max() < TOL_DB
    assert np.all(out[1:][g[1:] < -99.0] < -90.0)


def test_logmel_batch_strided_and_normalised():
    ys = np.stack([signals.hdr(100000, 10 + i) for i in range(5)])
    big = torch.zeros(5, 100100, device="cuda")
    big[:, :100000] = torch.from_numpy(ys).float().cuda()
    mean = np.linspace(-5, 5, 64).astype(np.float32)
    std = np.linspace(0.5, 2.0, 64).astype(np.float32)
    out = P.waveform_to_log_mel(big[:, :100000], mean=mean, std=std).cpu().numpy()
    ref = (R.waveform_to_log_mel(ys) - mean) / std          # SpectogramDataset.transform, logMel mode
    assert out.shape == (5, 7, 64)
    assert np.abs(out * std - ref * std).max() < TOL_DB


def test_stft_and_complex_to_logmel_dropins():
    y = signals.hdr(100000, 2)
    s = P.multichannel_stft(y[:, None])
    sr = R.multichannel_stft(y[:, None])
    assert s.shape == sr.shape == (1, 7, 16385) and s.dtype == np.complex64
    assert np.abs(s - sr).max() / np.abs(sr).max() < 1e-5
    lm3 = P.multichannel_complex_to_log_mel(sr)
    lm2 = P.multichannel_complex_to_log_mel(sr[0])          # 2-D input, Classical_methods/train_svm_detector.py:68
    ref = R.multichannel_complex_to_log_mel(sr)
    assert lm3.shape == (1, 7, 64) and lm2.shape == (7, 64) and lm3.dtype == np.float32
    assert np.abs(lm3 - ref).max() < TOL_DB and np.abs(lm2 - ref[0]).max() < TOL_DB
    two = np.stack([y, signals.white(100000, 4)], axis=1)    # two channels
    s2 = P.multichannel_stft(two)
    assert s2.shape == (2, 7, 16385)
    assert np.abs(s2[0] - sr[0]).max() / np.abs(sr).max() < 1e-5


def test_full_size_properties():
    """BASELINE-sized batch (16 x 60 s clips): size-independent properties instead of the (slow) oracle."""
    g = torch.Generator(device="cuda").manual_seed(0)
    w = (torch.randn(16, 2880000, device="cuda", generator=g) * 0.1).clamp_(-1, 1)
    a = P.waveform_to_log_mel(w)
    assert a.shape == (16, 182, 64) and bool(torch.isfinite(a).all())
    # power scaling: log-mel(2x) = log-mel(x) + 20 log10(2)
    b = P.waveform_to_log_mel(w * 2.0)
    assert float((b - a - 20 * np.log10(2.0)).abs().max()) < 2e-3
    # clip independence / determinism: a clip alone equals the clip inside the batch, bit for bit
    c = P.waveform_to_log_mel(w[5:6])
    assert torch.equal(c[0], a[5])
    # the oracle on two of the clips
    for i in (0, 15):
        ref = R.waveform_to_log_mel(w[i].cpu().numpy().astype(np.float64))
        assert np.abs(a[i].cpu().numpy() - ref).max() < TOL_DB


def test_error_behaviour():
    with pytest.raises(_ext.SedbError, match="reflect padding"):
        P.waveform_to_log_mel(torch.zeros(1, 16384, device="cuda"))
    assert P.waveform_to_log_mel(torch.zeros(0, 40000, device="cuda")).shape == (0, 3, 64)
    # all-zero audio sits on the amin floor
    assert float(P.waveform_to_log_mel(torch.zeros(1, 40000, device="cuda")).max()) == -100.0


def test_host_buffer_entry_point():
    lib = _ext.load()
    ys = np.stack([signals.hdr(60000, 20 + i) for i in range(3)]).astype(np.float32)
    wave = torch.from_numpy(ys).pin_memory()
    out = torch.empty(3, 4, 64).pin_memory()
    _ext.check(lib.sedb_logmel_host_f32(_ext.context(), ctypes.c_void_p(wave.data_ptr()), 3, 60000, 60000, None,
                                        ctypes.c_void_p(out.data_ptr())))
    assert np.abs(out.numpy() - R.waveform_to_log_mel(ys.astype(np.float64))).max() < TOL_DB


def _window_report(out, ref):
    """(worst |error| inside the dynamic-range window, worst |error| outside it, depth of the quietest bin below its
    frame's loudest bin, worst amount by which an outside bin is reported ABOVE its reference) in dB."""
    top = ref.max(axis=-1, keepdims=True)
    live = ref > top - DYN_RANGE_DB
    err = np.abs(out - ref)
    inside = float(err[live].max())
    outside = float(err[~live].max()) if (~live).any() else 0.0
    over = float((out - ref)[~live].max()) if (~live).any() else 0.0
    return inside, outside, float((top - ref).max()), over


def test_dynamic_range_contract_pcm16_tone_with_dither():
    """Worst realistic case for the contract: a full-scale 1 kHz tone quantised to 16-bit PCM with
    TPDF dither.  In float64 the quantisation-noise bins sit ~110-125 dB below the tone: some are OUTSIDE the window in
    which 1e-2 dB is guaranteed.  Inside: the 1e-2 dB bar.  Outside: the error is measured and bounded (a bin below the
    window may read high by the DFT's noise floor, never low by more), and the numbers are printed for DESIGN.md."""
    rng = np.random.default_rng(5)
    n = 480000
    t = np.arange(n) / 48000.0
    x = 0.999 * np.sin(2 * np.pi * 1000.0 * t)
    dither = (rng.random(n) - rng.random(n)) / 32768.0
    pcm = np.clip(np.round((x + dither) * 32768.0), -32768, 32767).astype(np.int16)
    y = pcm.astype(np.float64) / 32768.0                     # what soundfile.read hands the reference
    ref = R.waveform_to_log_mel(y)
    out = gpu_logmel(y)
    inside, outside, depth, over = _window_report(out, ref)
    print(f"pcm16 tone+dither: window {DYN_RANGE_DB:.0f} dB, deepest bin {depth:.1f} dB down, worst error inside "
          f"{inside:.2e} dB, outside {outside:.2e} dB (reads high by at most {over:.2e} dB)")
    assert inside < TOL_DB
    assert depth > DYN_RANGE_DB - 5.0                        # the case really reaches the edge of the window
    if DYN_RANGE_DB >= 100.0:
        assert outside < 1.0                                 # measured ~0.2 dB on B200 (fp16 split)
    else:                                                    # bf16 build: 45 dB of bins lie outside its 75 dB window; the
        top = ref.max(axis=-1, keepdims=True)                # contract there is only "stays below the window" (1.9 dB measured)
        assert np.all((out < top - DYN_RANGE_DB + 3.0)[ref <= top - DYN_RANGE_DB])
    # the same samples through the 16-bit PCM entry point are the same features
    from sed_b200.dataset.spectogram.preprocess import pcm16_to_log_mel
    out16 = pcm16_to_log_mel(torch.from_numpy(pcm[None]).cuda()).cpu().numpy()[0]
    assert _window_report(out16, ref)[0] < TOL_DB


def test_dynamic_range_contract_two_sources():
    """Loud 100 Hz hum + a 10 kHz component 90 dB below it (inside the 100 dB window of the fp16 build): the quiet
    source's mel bins must still be right to 1e-2 dB; with the component at -120 dB (outside) the error is reported."""
    n = 480000
    t = np.arange(n) / 48000.0
    for level_db, must_hold in ((-90.0, True), (-120.0, False)):
        y = 0.9 * np.sin(2 * np.pi * 100.0 * t) + 0.9 * 10 ** (level_db / 20) * np.sin(2 * np.pi * 10000.0 * t)
        ref = R.waveform_to_log_mel(y)
        out = gpu_logmel(y)
        k = int(np.argmax(ref[3, 40:])) + 40                 # the mel bin of the 10 kHz component
        err_quiet = float(np.abs(out[:, k] - ref[:, k])[1:-1].max())
        inside, outside, depth, over = _window_report(out, ref)
        print(f"two sources, quiet one at {level_db:.0f} dB: its mel bin {k} is {float((ref.max(-1) - ref[:, k])[1:-1].max()):.1f} dB "
              f"below the loudest bin, error there {err_quiet:.2e} dB; inside-window worst {inside:.2e}, outside {outside:.2e}")
        assert inside < TOL_DB
        if must_hold and DYN_RANGE_DB >= 100.0:
            assert err_quiet < TOL_DB


def _level_steps(n, levels, seed=0):
    """White noise whose level jumps between hop-aligned blocks (0 = digital silence)."""
    rng = np.random.default_rng(77 + seed)
    y = rng.standard_normal(n)
    hop = 15840
    for i in range(0, n, hop):
        y[i:i + hop] *= levels[(i // hop) % len(levels)]
    return np.clip(y, -1.0, 1.0)


@pytest.mark.parametrize("levels", [
    (1e-4, 1e-4, 0.3, 0.3, 1e-4),             # onsets of +70 dB: the provisional block scale is rejected and the frame refolded
    (0.0, 0.0, 0.2, 0.0, 1e-3, 0.25),         # digital silence on either side of sound
    (0.02, 0.09, 0.02, 0.005, 0.3),           # jumps across one and several scale steps (a step = 12 dB) in both directions
    (0.25, 1e-5, 1e-5, 1e-5),                 # a frame 88 dB below the previous one keeps its own scale
])
def test_block_scale_rule_on_level_steps(levels):
    """fp16 build: the per-frame block scale is first guessed from the half the frame shares with its predecessor
    (known before the frame is loaded); a second half that is louder by a scale step makes the kernel drop the stage-1
    attempt and fold the frame again (csrc/logmel.cuh, "block scale").  Parity must not depend on which way a frame went."""
    n = 15840 * 24 + 3000
    y = _level_steps(n, levels)
    out, ref = gpu_logmel(y), R.waveform_to_log_mel(y)
    assert out.shape == ref.shape
    assert_parity(out, ref)
    live = ref > -99.0                       # bins of silent frames sit at the amin floor of power_to_db (-100 dB)
    assert np.abs(out - ref)[live].max() < TOL_DB


def test_result_independent_of_batch_position():
    """A clip's log-mel must be bit-identical whatever batch it arrives in and wherever it sits in it: the frames are
    dealt to the SMs in consecutive runs whose boundaries depend on the batch, and a run's first frame finds its block
    scale by a pass of its own instead of from the previous frame; both ways must pick the same scale."""
    clip = _level_steps(15840 * 40 + 123, (0.05, 0.3, 0.3, 1e-3, 0.0, 0.1), seed=3)
    alone = gpu_logmel(clip[None])[0]
    rng = np.random.default_rng(5)
    for B, pos in [(3, 0), (3, 2), (7, 3), (150, 77)]:
        batch = (rng.standard_normal((B, clip.size)) * 0.1).astype(np.float32)
        batch[pos] = clip
        got = gpu_logmel(batch)[pos]
        assert np.array_equal(got, alone), (B, pos, np.abs(got - alone).max())
