"""ggml-experiments_b200: B200-native drop-in for the MobileViT forward path of datduonguva/ggml-experiments.

Python is only a thin ctypes mirror of the C ABI (include/ggml/ggml.h, include/mobilevit_b200.h) for tests and
bench.py; the product is the native libraries built by `build.py`.
"""
from .mobilevit import MobileViT, lib_ggml, lib_mobilevit, native_paths  # noqa: F401
