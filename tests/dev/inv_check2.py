import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
def steps(n, levels, seed=0):
    rng = np.random.default_rng(77 + seed); y = rng.standard_normal(n); hop = 15840
    for i in range(0, n, hop): y[i:i + hop] *= levels[(i // hop) % len(levels)]
    return np.clip(y, -1.0, 1.0)
clip = steps(15840 * 10 + 124, (0.05, 0.3, 0.3, 1e-3, 0.0, 0.1), seed=3).astype(np.float32)
buf = torch.zeros(clip.size + 8, device="cuda")
buf[:clip.size] = torch.from_numpy(clip).cuda()
print("ALIGNED", flush=True)
a = P.waveform_to_log_mel(buf[:clip.size][None]).cpu().numpy(); torch.cuda.synchronize()
buf[1:clip.size + 1] = torch.from_numpy(clip).cuda()
print("UNALIGNED", flush=True)
b = P.waveform_to_log_mel(buf[1:clip.size + 1][None]).cpu().numpy(); torch.cuda.synchronize()
print("diff per frame", np.abs(a - b).max(axis=-1))
