import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
wave = (torch.randn(64, 480000, device="cuda") * 0.1).clamp_(-1, 1)
for _ in range(5): P.waveform_to_log_mel(wave)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): P.waveform_to_log_mel(wave)
b.record(); torch.cuda.synchronize()
print(os.environ.get("SEDB_LIB_PATH", "default")[-12:], "logmel ms per 64 x 10 s:", a.elapsed_time(b) / 50)
