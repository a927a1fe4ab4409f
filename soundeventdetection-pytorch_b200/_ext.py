"""ctypes binding of libsedb.so (the C ABI declared in include/sedb.h).

There is deliberately no fallback: if the shared library is missing it is built with nvcc, and if that
fails (or no sm_100 device is present when a compute entry point is called) an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("SEDB_LIB_PATH") or os.path.join(_HERE, "libsedb.so")
_lock = threading.Lock()
_lib = None

c_float_p = ctypes.c_void_p      # raw device/host addresses are passed as integers
c_ll = ctypes.c_longlong

_SIGNATURES = {
    "sedb_version": (ctypes.c_int, []),
    "sedb_last_error": (ctypes.c_char_p, []),
    "sedb_split_is_fp16": (ctypes.c_int, []),
    "sedb_check_config": (ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.c_float] * 2),
    "sedb_num_frames": (c_ll, [c_ll]),
    "sedb_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p)]),
    "sedb_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sedb_mel_filterbank": (ctypes.c_int, [ctypes.c_void_p]),
    "sedb_logmel_f32": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_ll, c_ll, c_float_p, c_float_p,
                                       ctypes.c_void_p]),
    "sedb_stft_c64": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_ll, c_ll, c_float_p, ctypes.c_void_p]),
    "sedb_power_mel_db_f32": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_float_p, c_float_p,
                                             ctypes.c_void_p]),
    "sedb_logmel_host_f32": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_ll, c_ll, c_float_p, c_float_p]),
    "sedb_resample_num_samples": (c_ll, [c_ll, ctypes.c_int, ctypes.c_int]),
    "sedb_resample_filters": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 3),
    "sedb_resample_f32": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_ll, c_ll, ctypes.c_int, ctypes.c_int,
                                         c_float_p, c_ll, ctypes.c_void_p]),
    "sedb_logmel_pcm16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_ll, c_ll, c_ll, ctypes.c_int, c_float_p,
                                         c_float_p, ctypes.c_void_p]),
    "sedb_logmel_host_pcm16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_ll, c_ll, c_ll, ctypes.c_int,
                                              c_float_p, c_float_p]),
    "sedb_sed_host_pcm16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_ll, c_ll, c_ll,
                                           ctypes.c_int, c_float_p, c_float_p]),
    "sedb_cnn_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                       ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "sedb_cnn_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sedb_cnn_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                     ctypes.c_void_p]),
    "sedb_cnn_out_frames": (c_ll, [ctypes.c_void_p, c_ll]),
    "sedb_cnn_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p, c_ll, c_ll]),
    "sedb_cnn_forward": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_ll, c_float_p, c_float_p,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "sedb_cnn_workspace_invalidate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "sedb_m5_workspace_invalidate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "sedb_cnn_train_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p, c_ll, c_ll]),
    "sedb_cnn_train_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, c_float_p,
                                              c_ll, c_ll, ctypes.c_float, c_float_p, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_void_p]),
    "sedb_cnn_train_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, c_float_p,
                                               c_float_p, c_ll, c_ll, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "sedb_bce_with_logits": (ctypes.c_int, [c_float_p, c_float_p, c_ll, c_ll, c_ll, ctypes.c_int, ctypes.c_float,
                                            ctypes.c_float, c_float_p, c_float_p, ctypes.c_void_p]),
    "sedb_adam_amsgrad_step_dev": (ctypes.c_int, [c_float_p] * 5 + [c_ll, c_float_p, c_float_p] + [ctypes.c_float] * 5
                                   + [ctypes.c_void_p]),
    "sedb_m5_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "sedb_m5_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sedb_m5_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                    ctypes.c_void_p]),
    "sedb_m5_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p, c_ll]),
    "sedb_m5_forward": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_ll, c_float_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_void_p]),
    "sedb_adam_amsgrad_step": (ctypes.c_int, [c_float_p] * 5 + [c_ll] + [ctypes.c_float] * 5 + [c_ll, ctypes.c_float,
                                                                                                ctypes.c_void_p]),
    "sedb_sed_host_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_float_p, c_ll, c_ll, c_ll, c_float_p,
                                         c_float_p]),
    "sedb_debug_umma_probe": (ctypes.c_int, [c_float_p, c_float_p, c_float_p] + [ctypes.c_int] * 7
                              + [ctypes.c_void_p]),
    "sedb_debug_train_layout": (ctypes.c_int, [ctypes.c_void_p, c_ll, c_ll, ctypes.POINTER(c_ll), ctypes.c_int]),
    "sedb_debug_plan_layer": (ctypes.c_int, [ctypes.c_int] * 8 + [c_ll, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "sedb_debug_umma_rate": (ctypes.c_int, [ctypes.c_int] * 7 + [ctypes.c_void_p]),
    "sedb_debug_bulk_rate": (ctypes.c_int, [ctypes.c_int] * 7 + [ctypes.c_void_p]),
    "sedb_debug_phase_profile": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "sedb_launch_count": (c_ll, []),
}


class SedbError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load (building first if needed) libsedb.so and declare the prototypes of every exported symbol."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            from . import build as _build
            _build.build()
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == ABI drift; fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.sedb_version() != 1:
            raise SedbError(f"libsedb ABI version {lib.sedb_version()} != 1")
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        raise SedbError(load().sedb_last_error().decode("utf-8", "replace"))


_ctx = {}


def context():
    """Per-device sedb context (constant tables live on the device that is current at creation)."""
    import torch
    if not torch.cuda.is_available():
        raise SedbError("no CUDA device: the sedb hot path has no CPU fallback")
    dev = torch.cuda.current_device()
    h = _ctx.get(dev)
    if h is None:
        lib = load()
        out = ctypes.c_void_p()
        check(lib.sedb_create(ctypes.byref(out)))
        h = out
        _ctx[dev] = h
    return h


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
