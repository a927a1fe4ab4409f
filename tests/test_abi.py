"""The C-ABI shared library builds, loads and exports every symbol include/sedb.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

import sed_b200
from sed_b200 import _ext
from oracle import logmel_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "sedb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sedb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ext.lib_path()) if os.path.exists(_ext.lib_path()) else _ext.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sedb.h but not exported by libsedb.so"


def test_binding_covers_header():
    assert sorted(_ext._SIGNATURES.keys()) == header_functions()


def test_version_and_config_check():
    lib = _ext.load()
    assert lib.sedb_version() == 1
    assert lib.sedb_check_config(48000, 31680, 15840, 32768, 64, 20.0, 24000.0) == 0
    assert lib.sedb_check_config(44100, 31680, 15840, 32768, 64, 20.0, 24000.0) != 0
    assert b"configuration mismatch" in lib.sedb_last_error()
    with pytest.raises(_ext.SedbError):
        _ext.check(lib.sedb_check_config(48000, 31680, 15840, 32768, 128, 20.0, 24000.0))


@pytest.mark.parametrize("n", [31680, 480000, 2880000, 2880001])
def test_num_frames(n):
    assert _ext.load().sedb_num_frames(n) == logmel_ref.num_frames(n)


def test_mel_filterbank_matches_oracle_bit_for_bit():
    out = np.empty((16385, 64), dtype=np.float32)
    assert _ext.load().sedb_mel_filterbank(ctypes.c_void_p(out.ctypes.data)) == 0
    assert np.array_equal(out, logmel_ref.mel_filter_bank_matrix())


def test_null_arguments_are_rejected_not_crashing():
    lib = _ext.load()
    assert lib.sedb_mel_filterbank(None) != 0
    assert lib.sedb_create(None) != 0
    assert lib.sedb_logmel_f32(None, None, 1, 100000, 100000, None, None, None) != 0
    assert lib.sedb_cnn_forward(None, None, 1, 30, None, None, None, 0, None) != 0
    assert lib.sedb_m5_forward(None, None, 1, None, None, 0, None) != 0
    assert lib.sedb_destroy(None) == 0 and lib.sedb_cnn_destroy(None) == 0 and lib.sedb_m5_destroy(None) == 0
