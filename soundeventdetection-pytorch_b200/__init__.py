"""B200-native hot path of ariel415el/SoundEventDetection-Pytorch (import as ``sed_b200``).

Mirrors the reference's module layout for the path it accelerates:

* ``sed_b200.dataset.spectogram.preprocess``   <-> reference ``dataset/spectogram/preprocess.py``
* ``sed_b200.models.spectogram_models``        <-> reference ``models/spectogram_models.py``
* ``sed_b200.models.waveform_models``          <-> reference ``models/waveform_models.py``

All compute goes through ``libsedb.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/sedb.h``).
There is no CPU fallback.
"""
__version__ = "0.1.0"
