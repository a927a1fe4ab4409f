import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200 import _ext
from sed_b200.dataset.spectogram import preprocess as P
lib = _ext.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = (torch.randn(B, 2880000, device="cuda") * 0.1).clamp_(-1, 1)
if len(sys.argv) > 2 and sys.argv[2] == "pcm16":
    w = (w * 32767.0).round_().to(torch.int16)
for _ in range(2): P.waveform_to_log_mel(w)
torch.cuda.synchronize()
lib.sedb_debug_phase_profile(1, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); P.waveform_to_log_mel(w); e1.record(); torch.cuda.synchronize()
out = np.zeros(128, dtype=np.uint64)
lib.sedb_debug_phase_profile(0, ctypes.c_void_p(out.ctypes.data))
frames = B * 182
names = ["load+scale", "fold/split", "wait S1", "twiddle/radix2", "wait S2", "power", "row128", "final sync", "sync->mel", "mel partials", "sync", "finalize", "aux 12", "aux 13", "aux 14"]
tot = out[:12].sum()
print(f"B={B} {e0.elapsed_time(e1):.3f} ms; cycles/frame total {tot/frames:.0f}")
for n, v in zip(names, out[:15]): print(f"  {n:16s} {v/frames:8.0f} cyc/frame  {100*v/tot:5.1f}%")
