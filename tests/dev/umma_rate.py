import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sed_b200
from sed_b200 import _ext
lib = _ext.load(); torch.zeros(1, device="cuda")
out = np.zeros(1, dtype=np.uint64)
reps = 512
for mode in (0, 1):
  for grid in (1, 148):
    for N in (32, 64, 128):
        for b_major in (0, 1):
            for n_acc in (1, 4):
                la, lb = 2048, N * 16
                rc = lib.sedb_debug_umma_rate(N, b_major, n_acc, reps, la, lb | (mode << 24), grid, ctypes.c_void_p(out.ctypes.data))
                print(f"mode {mode} grid {grid:3d} N {N:3d} b_major {b_major} n_acc {n_acc}: {out[0]/reps:7.1f} cyc/MMA (rc {rc})")
