"""ctypes binding of the CPU oracle (oracle/mobilevit_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmvit_oracle.so")
PURE_F32 = 1
NO_ACT_ROUND = 2
LEGACY_F16_TABLES = 4  # SiLU / softmax-exp through f16 lookup tables (the ggml the author ran)

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mobilevit_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.mvo_load.restype = ctypes.c_void_p
        L.mvo_load.argtypes = [ctypes.c_char_p]
        L.mvo_free.argtypes = [ctypes.c_void_p]
        L.mvo_num_tensors.argtypes = [ctypes.c_void_p]
        L.mvo_out_channels.argtypes = [ctypes.c_void_p]
        L.mvo_num_weights.argtypes = [ctypes.c_void_p]
        L.mvo_num_weights.restype = ctypes.c_long
        L.mvo_forward.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, _f32p, _f32p, ctypes.c_int,
                                  ctypes.POINTER(_f32p)]
        L.mvo_forward_batch.restype = ctypes.c_double
        L.mvo_forward_batch.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p,
                                        _f32p, ctypes.c_int, ctypes.c_int]
        L.mvo_gemm.argtypes = [ctypes.c_int] * 3 + [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int]
        L.mvo_conv2d.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int]
        L.mvo_dwconv2d.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, _f32p, ctypes.c_int]
        L.mvo_unfold.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        L.mvo_fold.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        L.mvo_layernorm.argtypes = [_f32p, ctypes.c_int, ctypes.c_size_t, _f32p, _f32p, ctypes.c_float, _f32p]
        L.mvo_softmax_rows.argtypes = [_f32p, ctypes.c_int, ctypes.c_size_t]
        L.mvo_max_threads.restype = ctypes.c_int
        L.mvo_num_classes.argtypes = [ctypes.c_void_p]
        L.mvo_classify.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, _f32p]
        L.mvo_preprocess_u8.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


class OracleModel:
    """mobilevit_model + extract_features (main.cpp:202-213,604-646) on the CPU oracle."""

    def __init__(self, weight_path: str):
        self._h = lib().mvo_load(weight_path.encode())
        if not self._h:
            raise FileNotFoundError(weight_path)
        self.out_channels = lib().mvo_out_channels(self._h)
        self.num_tensors = lib().mvo_num_tensors(self._h)
        self.num_weights = lib().mvo_num_weights(self._h)

    def close(self):
        if self._h:
            lib().mvo_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def forward(self, imgs_hwc: np.ndarray, flags: int = 0, n_threads: int = 0, return_time: bool = False):
        """imgs_hwc [N,H,W,3] f32 -> (features [N,C,H/32,W/32], pooled [N,C])."""
        imgs = np.ascontiguousarray(imgs_hwc, dtype=np.float32)
        n, h, w, _ = imgs.shape
        feat = np.empty((n, self.out_channels, h // 32, w // 32), dtype=np.float32)
        pooled = np.empty((n, self.out_channels), dtype=np.float32)
        secs = lib().mvo_forward_batch(self._h, _p(imgs), n, h, w, _p(feat), _p(pooled), flags, n_threads)
        if return_time:
            return feat, pooled, secs
        return feat, pooled

    @property
    def num_classes(self) -> int:
        return lib().mvo_num_classes(self._h)

    def classify(self, pooled: np.ndarray) -> np.ndarray:
        """Classifier head on pooled features [N,C] -> logits [N,classes]."""
        pooled = np.ascontiguousarray(pooled, dtype=np.float32)
        logits = np.empty((pooled.shape[0], self.num_classes), dtype=np.float32)
        rc = lib().mvo_classify(self._h, _p(pooled), pooled.shape[0], _p(logits))
        if rc:
            raise ValueError("weight file has no classifier")
        return logits

    def forward_stages(self, img_hwc: np.ndarray, stage_shapes, flags: int = 0):
        """One image; returns the 7 stage outputs (stem, layer1..5, exp) as CHW arrays."""
        img = np.ascontiguousarray(img_hwc, dtype=np.float32)
        h, w, _ = img.shape
        bufs = [np.empty(s, dtype=np.float32) for s in stage_shapes]
        arr = (_f32p * 7)(*[_p(b) for b in bufs])
        lib().mvo_forward(self._h, _p(img), h, w, None, None, flags, arr)
        return bufs


def preprocess_u8(images: np.ndarray, h: int, w: int) -> np.ndarray:
    """sam_image_preprocess (main.cpp:538-601) per image: [N,src_h,src_w,3] u8 -> [N,h,w,3] f32 (stride-W fix, App. C #3)."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, sh, sw, _ = images.shape
    out = np.empty((n, h, w, 3), dtype=np.float32)
    for i in range(n):
        lib().mvo_preprocess_u8(images[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), sh, sw, h, w, _p(out[i]))
    return out
