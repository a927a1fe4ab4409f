"""tcgen05.mma rate when consecutive MMAs switch B operand / accumulator (the conv kernels' issue pattern)."""
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
    for n_acc in (1, 4):
        for commit_log2 in (0, 2, 3, 4, 6):
            lb = (8 << 12) | (1 << 16) | (2 << 24)
            rc = lib.sedb_debug_umma_rate(N, 0, n_acc, reps, 2048 | (commit_log2 << 20), lb, 148, ctypes.c_void_p(out.ctypes.data))
            every = (1 << commit_log2) if commit_log2 else 0
            print(f"N {N:3d} n_acc {n_acc} commit every {every:3d} MMAs: {out[0]/reps:7.1f} cyc/MMA (rc {rc})", flush=True)
