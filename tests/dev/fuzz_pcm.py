"""Random level steps / channel counts through the 16-bit PCM log-mel entry point vs the CPU oracle (development fuzz)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
from oracle import logmel_ref as R, audio_ref as A
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for trial in range(10):
    B = int(rng.integers(1, 24)); C = int(rng.choice([1, 2, 3, 4, 6])); n = int(rng.integers(16385, 15840 * 12))
    y = rng.standard_normal((B, n, C))
    for b in range(B):
        for i in range(0, n, 15840):
            y[b, i:i + 15840] *= (10.0 ** rng.uniform(-4, 0) if rng.random() > 0.15 else 0.0)
    pcm = np.clip(np.round(y * 32767), -32768, 32767).astype(np.int16)
    out = P.pcm16_to_log_mel(torch.from_numpy(pcm).cuda()).cpu().numpy()
    ref = np.stack([R.waveform_to_log_mel(A.pcm16_to_mono(pcm[b])) for b in range(B)])
    top = ref.max(axis=-1, keepdims=True)
    live = (ref > top - 100.0) & (ref > -99.0)
    err = np.abs(out - ref)[live].max() if live.any() else 0.0
    worst = max(worst, err)
    print(trial, B, C, n, "max err in window", err, flush=True)
    assert err < 1e-2
print("pcm fuzz ok, worst", worst)
