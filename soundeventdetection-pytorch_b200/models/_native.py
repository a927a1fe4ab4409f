"""Shared plumbing of the drop-in nn.Modules: native handle lifetime, parameter repacking, workspaces."""
from __future__ import annotations

import ctypes

import torch

from .. import _ext


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class NativeHandle:
    """Owns a sedb_cnn_t / sedb_m5_t for one CUDA device and keeps it in sync with the module's tensors."""

    def __init__(self, create, destroy, load, invalidate=None):
        self._create, self._destroy, self._load, self._invalidate = create, destroy, load, invalidate
        self._handles = {}       # device index -> (handle, parameter fingerprint)
        self._workspaces = {}    # (device, key) -> uint8 tensor

    def mark_dirty(self):
        """Forget the parameter fingerprints: the next get() repacks.  Called when the module enters or leaves train mode
        (a torch train-mode forward updates BatchNorm running statistics without bumping their version counter)."""
        for entry in self._handles.values():
            entry[1] = None

    def get(self, device: torch.device, tensors):
        """Returns the handle for `device`, (re)loading parameters when any tensor changed since the last call."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        fp = tuple((t.data_ptr(), t._version) for t in tensors)
        entry = self._handles.get(idx)
        if entry is None:
            h = ctypes.c_void_p()
            self._create(ctypes.byref(h))
            entry = [h, None]
            self._handles[idx] = entry
        if entry[1] != fp:
            prepared = [t.detach().to(device=device, dtype=torch.float32).contiguous() for t in tensors]
            arr = (ctypes.c_void_p * len(prepared))(*[t.data_ptr() for t in prepared])
            self._load(entry[0], arr, len(prepared))
            torch.cuda.current_stream(device).synchronize()      # `prepared` temporaries may be freed after this
            entry[1] = fp
        return entry[0]

    def workspace(self, device: torch.device, key, nbytes: int) -> torch.Tensor:
        k = (device.index, key)
        ws = self._workspaces.get(k)
        if ws is None or ws.numel() < nbytes:
            kind = key[0] == "train" if isinstance(key, tuple) and key else False
            # keep at most one live workspace per kind (inference / training) and module
            for old in [q for q in self._workspaces if (isinstance(q[1], tuple) and q[1] and q[1][0] == "train") == kind]:
                del self._workspaces[old]
            # the library remembers which (pointer, geometry) it has zeroed the padding of; memory handed back by the
            # caching allocator may have been scribbled on since, so forget everything it knew
            if self._invalidate is not None:
                for h, _ in self._handles.values():
                    self._invalidate(h, None)
            ws = torch.empty(nbytes + 128, dtype=torch.uint8, device=device)
            self._workspaces[k] = ws
        return ws

    def close(self):
        for h, _ in self._handles.values():
            try:
                self._destroy(h)
            except Exception:
                pass
        self._handles.clear()
        self._workspaces.clear()

    def __del__(self):
        self.close()

    # a handle is device state of ONE module object: copies and pickles of the module start without it and rebuild it
    # lazily on their first forward (ctypes pointers cannot be pickled, and two owners would double-free)
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (type(None), ())


def aligned_ptr(ws: torch.Tensor):
    """128-byte aligned address inside a uint8 workspace tensor and the bytes left behind it."""
    base = ws.data_ptr()
    off = (-base) % 128
    return ctypes.c_void_p(base + off), ws.numel() - off
