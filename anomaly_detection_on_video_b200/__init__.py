"""B200-native (sm_100a) implementation of the I3D snippet-feature extraction hot path of
jinmang2/anomaly_detection_on_video.

Python is the API surface (drop-ins for the reference's ``src/i3d.py``, ``src/dataset.py`` and
``extract_features.py``); all compute goes through the C ABI of ``libvad_b200.so``
(``include/vad_b200.h``), hand-written CUDA for sm_100a.  There is no CPU fallback.
"""
__version__ = "0.1.0"

__all__ = ["__version__"]
