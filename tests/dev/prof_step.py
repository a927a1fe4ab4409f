"""Small driver for ncu captures: a few passes of the hot path (log-mel + CNN) over a small batch."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200.dataset.spectogram import preprocess as P
import refmodels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = (torch.randn(B, 2880000, device="cuda") * 0.1).clamp_(-1, 1)
m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda()
mean = torch.full((64,), 18.0, device="cuda"); std = torch.full((64,), 6.0, device="cuda")
for _ in range(reps):
    x = P.waveform_to_log_mel(w, mean=mean, std=std)
    p = m.logits(x[:, None])
torch.cuda.synchronize()
print("done", p.shape)
