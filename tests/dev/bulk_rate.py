"""cp.async.bulk (1-D TMA) global->shared throughput per SM vs copy size, copies in flight, issuing warps and CTAs
(L2-resident source)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import sed_b200
from sed_b200 import _ext
lib = _ext.load()
_ext.context()
out = ctypes.c_ulonglong()
for grid in (148, 8):
    for bytes_, depth, nw in ((4096, 4, 1), (4096, 4, 2), (4096, 4, 4), (4096, 4, 8), (8192, 2, 8), (24576, 2, 1), (24576, 2, 2),
                              (24576, 2, 4), (65536, 2, 1), (98304, 2, 1), (49152, 2, 2), (1024, 4, 8), (16384, 3, 4)):
        reps = 480
        _ext.check(lib.sedb_debug_bulk_rate(bytes_, depth, reps, 24, 0, grid, nw, ctypes.byref(out)))
        print(f"grid {grid:3d} copy {bytes_:6d} B depth {depth:2d} warps {nw}: {out.value / reps:8.0f} cyc/copy "
              f"{bytes_ * reps / out.value:6.1f} B/clk/SM", flush=True)
