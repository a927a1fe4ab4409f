import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200
from sed_b200.dataset import dataset_utils as DU
for orig, C in ((44100, 64), (16000, 64), (96000, 64)):
    x = torch.randn(C, orig * 60, device="cuda") * 0.1
    for _ in range(2): DU.resample(x, orig, 48000)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): y = DU.resample(x, orig, 48000)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{orig} -> 48000, {C} clips x 60 s: {ms:.3f} ms ({C * 60 / 3600 / (ms * 1e-3):.0f} audio-hours/s)")
