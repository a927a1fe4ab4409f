import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sed_b200, signals
from sed_b200.dataset.spectogram import preprocess as P
from oracle import logmel_ref
ys = np.stack([signals.hdr(2880000, 30 + i) for i in range(2)])
lm = logmel_ref.waveform_to_log_mel(ys)
for trial in range(3):
    g = P.waveform_to_log_mel(torch.from_numpy(ys).float().cuda()).cpu().numpy()
    d = np.abs(g - lm).max(axis=-1)
    print("trial", trial, "max err", d.max(), "bad frames", [(int(c), int(t), float(d[c, t])) for c, t in zip(*np.nonzero(d > 1e-2))][:20])
g1 = np.stack([P.waveform_to_log_mel(torch.from_numpy(ys[i:i+1]).float().cuda()).cpu().numpy()[0] for i in range(2)])
d = np.abs(g1 - lm).max(axis=-1)
print("single", d.max(), [(int(c), int(t), float(d[c, t])) for c, t in zip(*np.nonzero(d > 1e-2))][:20])
print("batch vs single identical:", np.array_equal(g, g1), np.abs(g-g1).max())
