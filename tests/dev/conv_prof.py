import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200 import _ext
import refmodels
lib = _ext.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda()
x = torch.randn(B, 1, 182, 64, device="cuda")
for _ in range(2): m.logits(x)
torch.cuda.synchronize()
lib.sedb_debug_phase_profile(1, None)
m.logits(x); torch.cuda.synchronize()
out = np.zeros(128, dtype=np.uint64)
lib.sedb_debug_phase_profile(0, ctypes.c_void_p(out.ctypes.data))
names = ["epi wait", "epi work", "mma wait patch", "mma issue", "mma wait tmem", "copy wait patch_free", "items", "mma wait W"]
for L in range(7):
    c = out[16 * (L + 1): 16 * (L + 2)]
    items = max(1, int(c[6]))
    print(f"layer {L}: items/CTA-thread0 {items}", " | ".join(f"{n} {int(v)//items}" for n, v in zip(names, c[:8])))
