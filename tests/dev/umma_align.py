import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sed_b200
from sed_b200 import _ext
lib = _ext.load(); torch.zeros(1, device="cuda")
out = np.zeros(1, dtype=np.uint64)
reps = 512
for N in (32, 64, 128):
    for la in (2048, 8448):
        for off in (0, 1, 2, 4, 66):
            rc = lib.sedb_debug_umma_rate(N, 0, 4, reps, la | (off << 24), (N * 16) | (1 << 24), 148, ctypes.c_void_p(out.ctypes.data))
            print(f"N {N:3d} lbo_a {la:5d} A offset {off*16:5d} B: {out[0]/reps:7.1f} cyc/MMA (rc {rc})")
