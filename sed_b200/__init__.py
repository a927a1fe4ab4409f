"""Importable alias of the ``soundeventdetection-pytorch_b200`` package directory.

The product package lives in a directory whose name (mandated by the project layout) contains a hyphen and
therefore cannot be imported by name.  ``import sed_b200`` makes that directory importable: this package's
``__path__`` points at it and its ``__init__`` is executed in this namespace, so
``sed_b200.models.spectogram_models`` resolves to
``soundeventdetection-pytorch_b200/models/spectogram_models.py``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "soundeventdetection-pytorch_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
