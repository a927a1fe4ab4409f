import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
def steps(n, levels, seed=0):
    rng = np.random.default_rng(77 + seed); y = rng.standard_normal(n); hop = 15840
    for i in range(0, n, hop): y[i:i + hop] *= levels[(i // hop) % len(levels)]
    return np.clip(y, -1.0, 1.0)
def lm(y): return P.waveform_to_log_mel(torch.from_numpy(np.asarray(y, dtype=np.float32)).cuda()).cpu().numpy()
for extra in (123, 124, 0):
    clip = steps(15840 * 40 + extra, (0.05, 0.3, 0.3, 1e-3, 0.0, 0.1), seed=3)
    alone = lm(clip[None])[0]
    rng = np.random.default_rng(5)
    for B, pos in [(3, 0), (3, 1), (3, 2), (7, 3), (150, 77), (150, 78)]:
        batch = (rng.standard_normal((B, clip.size)) * 0.1).astype(np.float32); batch[pos] = clip
        got = lm(batch)[pos]
        d = np.abs(got - alone).max(axis=1)
        print(os.environ.get("SEDB_LIB_PATH", "new")[-12:], extra, B, pos, "max", d.max(), "frames differing", np.nonzero(d)[0][:12])
