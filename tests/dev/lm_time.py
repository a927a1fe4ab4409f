import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(1)
wave = torch.empty(C, 2880000, device="cuda")
for i in range(C): wave[i] = (torch.randn(2880000, device="cuda", generator=g) * 0.1).clamp_(-1, 1)
for _ in range(3): P.waveform_to_log_mel(wave)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): P.waveform_to_log_mel(wave)
b.record(); torch.cuda.synchronize()
print(os.environ.get("SEDB_LIB_PATH", "default"), "logmel ms per", C, "clips:", a.elapsed_time(b) / 10)
