"""tcgen05.mma rate vs the K-group strides (LBO) of the A and B shared-memory operands (no-swizzle K-major)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sed_b200
from sed_b200 import _ext
lib = _ext.load(); torch.zeros(1, device="cuda")
out = np.zeros(1, dtype=np.uint64)
reps = 512
for N in (64, 128):
    for la in (2048, 4704, 2368, 4192):
        for lb in (N * 16, N * 32, N * 32 + 128, 4096, 4096 + 16):
            for n_acc in (2, 4):
                rc = lib.sedb_debug_umma_rate(N, 0, n_acc, reps, la, lb | (1 << 24), 148, ctypes.c_void_p(out.ctypes.data))
                print(f"N {N:3d} lbo_a {la:5d} lbo_b {lb:5d} n_acc {n_acc}: {out[0]/reps:7.1f} cyc/MMA (rc {rc})", flush=True)
