import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
C = 256
g = torch.Generator(device="cuda").manual_seed(1)
wave = torch.empty(C, 2880000, device="cuda")
for i in range(C): wave[i] = (torch.randn(2880000, device="cuda", generator=g) * 0.1).clamp_(-1, 1)
mean = torch.full((64,), 18.0, device="cuda"); std = torch.full((64,), 6.0, device="cuda")
for kw in ({}, {"mean": mean, "std": std}, {}):
    for _ in range(3): P.waveform_to_log_mel(wave, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): P.waveform_to_log_mel(wave, **kw)
    b.record(); torch.cuda.synchronize()
    print("norm" if kw else "plain", a.elapsed_time(b) / 10)
