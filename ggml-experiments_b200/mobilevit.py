"""ctypes mirror of include/mobilevit_b200.h -- the host-side equivalent of the reference's
`mobilevit_model` (load_model_v2 + extract_features, /root/reference/mobilevit/main.cpp:314-515,604-646).

There is no CPU implementation behind these calls: without the native CUDA libraries the import of the
shared objects fails loudly, and without a GPU the first compute call aborts inside libggml_b200."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_f32p = ctypes.POINTER(ctypes.c_float)


def native_paths() -> dict:
    return {"ggml": os.path.join(_BUILD, "libggml_b200.so"), "mobilevit": os.path.join(_BUILD, "libmobilevit_b200.so")}


class PlanInfo(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("graph_nodes", ctypes.c_int), ("launches", ctypes.c_int),
                ("arena_bytes", ctypes.c_int64), ("naive_bytes", ctypes.c_int64), ("weight_bytes", ctypes.c_int64),
                ("cuda_graph", ctypes.c_int), ("lanes", ctypes.c_int)]


_ggml = None
_mv = None


def lib_ggml() -> ctypes.CDLL:
    global _ggml
    if _ggml is None:
        p = native_paths()["ggml"]
        if not os.path.exists(p):
            raise ImportError(f"{p} not built: run `python ggml-experiments_b200/build.py` (needs nvcc). "
                              "There is no CPU/PyTorch fallback.")
        L = ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        L.ggml_b200_set_mode.argtypes = [ctypes.c_int]
        L.ggml_b200_get_mode.restype = ctypes.c_int
        L.ggml_b200_device_count.restype = ctypes.c_int
        L.ggml_b200_set_device.argtypes = [ctypes.c_int]
        L.ggml_b200_set_stream.argtypes = [ctypes.c_void_p]
        L.ggml_b200_get_stream.restype = ctypes.c_void_p
        L.ggml_b200_version.restype = ctypes.c_char_p
        L.ggml_fp32_to_fp16.argtypes = [ctypes.c_float]
        L.ggml_fp32_to_fp16.restype = ctypes.c_uint16
        L.ggml_fp16_to_fp32.argtypes = [ctypes.c_uint16]
        L.ggml_fp16_to_fp32.restype = ctypes.c_float
        _ggml = L
    return _ggml


def lib_mobilevit() -> ctypes.CDLL:
    global _mv
    if _mv is None:
        lib_ggml()
        p = native_paths()["mobilevit"]
        if not os.path.exists(p):
            raise ImportError(f"{p} not built: run `python ggml-experiments_b200/build.py`.")
        L = ctypes.CDLL(p)
        L.mvit_load.restype = ctypes.c_void_p
        L.mvit_load.argtypes = [ctypes.c_char_p]
        L.mvit_free.argtypes = [ctypes.c_void_p]
        L.mvit_num_tensors.argtypes = [ctypes.c_void_p]
        L.mvit_num_weights.argtypes = [ctypes.c_void_p]
        L.mvit_num_weights.restype = ctypes.c_int64
        L.mvit_out_channels.argtypes = [ctypes.c_void_p]
        L.mvit_extract_features.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p]
        for name in ("mvit_prepare", "mvit_forward_device"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        for name in ("mvit_device_input", "mvit_device_features", "mvit_device_pooled"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            getattr(L, name).restype = ctypes.c_void_p
        for name in ("mvit_host_input", "mvit_host_features", "mvit_host_pooled"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            getattr(L, name).restype = _f32p
        L.mvit_compute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.mvit_profile_json.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_char_p, ctypes.c_size_t]
        L.mvit_debug_stage.restype = ctypes.c_int64
        L.mvit_debug_stage.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p,
                                       ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        for name in ("mvit_slot_input", "mvit_slot_features", "mvit_slot_pooled"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            getattr(L, name).restype = _f32p
        L.mvit_slot_set_transfers.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6
        for name in ("mvit_slot_submit", "mvit_slot_wait"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.mvit_release.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.mvit_num_classes.argtypes = [ctypes.c_void_p]
        L.mvit_classify.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.POINTER(ctypes.c_int32)]
        L.mvit_host_logits.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.mvit_host_logits.restype = _f32p
        L.mvit_slot_logits.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4
        L.mvit_slot_logits.restype = _f32p
        L.mvit_host_input_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5
        L.mvit_host_input_u8.restype = u8p
        L.mvit_compute_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5
        L.mvit_slot_input_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6
        L.mvit_slot_input_u8.restype = u8p
        L.mvit_slot_submit_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6
        L.mvit_preprocess_u8.argtypes = [ctypes.c_void_p, u8p] + [ctypes.c_int] * 5 + [_f32p]
        L.mvit_plan_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PlanInfo)]
        _mv = L
    return _mv


FAST, EXACT, EXACT_F32 = 0, 1, 2


def set_mode(mode: int) -> None:
    lib_ggml().ggml_b200_set_mode(mode)


class MobileViT:
    """`mobilevit_model` of the reference: load a weight file, extract features for a batch of HWC images."""

    def __init__(self, weight_path: str):
        self._L = lib_mobilevit()
        self._h = self._L.mvit_load(os.fsencode(weight_path))
        if not self._h:
            raise FileNotFoundError(f"cannot load weight file {weight_path}")
        self.out_channels = self._L.mvit_out_channels(self._h)
        self.num_tensors = self._L.mvit_num_tensors(self._h)
        self.num_weights = self._L.mvit_num_weights(self._h)
        self.num_classes = self._L.mvit_num_classes(self._h)  # 0: the file has no classifier head

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.mvit_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def extract_features(self, images_hwc: np.ndarray):
        """images [N,H,W,3] f32 in [0,1] -> (features [N,C,H/32,W/32], pooled [N,C]); host in, host out."""
        imgs = np.ascontiguousarray(images_hwc, dtype=np.float32)
        n, h, w, c = imgs.shape
        assert c == 3
        feat = np.empty((n, self.out_channels, h // 32, w // 32), dtype=np.float32)
        pooled = np.empty((n, self.out_channels), dtype=np.float32)
        rc = self._L.mvit_extract_features(self._h, imgs.ctypes.data_as(_f32p), n, h, w, feat.ctypes.data_as(_f32p),
                                           pooled.ctypes.data_as(_f32p))
        if rc != 0:
            raise ValueError(f"mvit_extract_features: invalid arguments n={n} h={h} w={w}")
        return feat, pooled

    def classify(self, images_hwc: np.ndarray):
        """images [N,H,W,3] f32 -> (logits [N,classes], top-1 ids [N]); needs a weight file with a classifier head."""
        imgs = np.ascontiguousarray(images_hwc, dtype=np.float32)
        n, h, w, _ = imgs.shape
        logits = np.empty((n, max(self.num_classes, 1)), dtype=np.float32)
        top1 = np.empty(n, dtype=np.int32)
        rc = self._L.mvit_classify(self._h, imgs.ctypes.data_as(_f32p), n, h, w, logits.ctypes.data_as(_f32p),
                                   top1.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        if rc != 0:
            raise ValueError("mvit_classify: " + ("weight file has no classifier" if rc == 2 else "invalid arguments"))
        return logits, top1

    def host_logits(self, n, h, w) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.mvit_host_logits(self._h, n, h, w), shape=(n, self.num_classes))

    def slot_logits(self, n, h, w, slot) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.mvit_slot_logits(self._h, n, h, w, slot), shape=(n, self.num_classes))

    # ---- u8 images, preprocessing on the device (sam_image_preprocess, main.cpp:538-601) ----
    def preprocess_u8(self, images_u8: np.ndarray, h: int, w: int) -> np.ndarray:
        imgs = np.ascontiguousarray(images_u8, dtype=np.uint8)
        n, sh, sw, _ = imgs.shape
        out = np.empty((n, h, w, 3), dtype=np.float32)
        if self._L.mvit_preprocess_u8(self._h, imgs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), n, sh, sw, h, w, out.ctypes.data_as(_f32p)) != 0:
            raise ValueError("mvit_preprocess_u8: invalid arguments")
        return out

    def host_input_u8(self, n, h, w, src_h, src_w) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.mvit_host_input_u8(self._h, n, h, w, src_h, src_w), shape=(n, src_h, src_w, 3))

    def compute_u8(self, n, h, w, src_h, src_w):
        if self._L.mvit_compute_u8(self._h, n, h, w, src_h, src_w) != 0:
            raise ValueError("mvit_compute_u8: invalid arguments")
        c = self.out_channels
        f = np.ctypeslib.as_array(self._L.mvit_host_features(self._h, n, h, w), shape=(n, c, h // 32, w // 32))
        p = np.ctypeslib.as_array(self._L.mvit_host_pooled(self._h, n, h, w), shape=(n, c))
        return f, p

    def slot_input_u8(self, n, h, w, slot, src_h, src_w) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.mvit_slot_input_u8(self._h, n, h, w, slot, src_h, src_w), shape=(n, src_h, src_w, 3))

    def slot_submit_u8(self, n, h, w, slot, src_h, src_w) -> None:
        if self._L.mvit_slot_submit_u8(self._h, n, h, w, slot, src_h, src_w) != 0:
            raise ValueError("mvit_slot_submit_u8: invalid arguments")

    # ---- zero-copy host interface: write the pinned input buffer in place, compute, read the outputs ----
    def host_input(self, n, h, w) -> np.ndarray:
        p = self._L.mvit_host_input(self._h, n, h, w)
        return np.ctypeslib.as_array(p, shape=(n, h, w, 3))

    def compute(self, n, h, w):
        if self._L.mvit_compute(self._h, n, h, w) != 0:
            raise ValueError("mvit_compute: invalid shape")
        c = self.out_channels
        f = np.ctypeslib.as_array(self._L.mvit_host_features(self._h, n, h, w), shape=(n, c, h // 32, w // 32))
        p = np.ctypeslib.as_array(self._L.mvit_host_pooled(self._h, n, h, w), shape=(n, c))
        return f, p

    def debug_stage(self, n, h, w, idx) -> np.ndarray:
        """Stage tap idx as [N,C,H,W] (needs MVIT_DEBUG_STAGES=1 in the environment before the shape is first used)."""
        ne = (ctypes.c_int64 * 4)()
        cap = n * h * w * 16
        buf = np.empty(cap, dtype=np.float32)
        cnt = self._L.mvit_debug_stage(self._h, n, h, w, idx, buf.ctypes.data_as(_f32p), cap, ne)
        if cnt < 0:
            raise RuntimeError(f"mvit_debug_stage rc={cnt}")
        return buf[:cnt].reshape(ne[3], ne[2], ne[1], ne[0]).copy()

    # ---- pipelined slots (throughput serving) ----
    def slot_input(self, n, h, w, slot) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.mvit_slot_input(self._h, n, h, w, slot), shape=(n, h, w, 3))

    def slot_submit(self, n, h, w, slot) -> None:
        if self._L.mvit_slot_submit(self._h, n, h, w, slot) != 0:
            raise ValueError("mvit_slot_submit: invalid arguments")

    def slot_set_transfers(self, n, h, w, slot, upload: bool, download: bool) -> None:
        self._L.mvit_slot_set_transfers(self._h, n, h, w, slot, int(upload), int(download))

    def slot_wait(self, n, h, w, slot):
        if self._L.mvit_slot_wait(self._h, n, h, w, slot) != 0:
            raise ValueError("mvit_slot_wait: invalid arguments")
        c = self.out_channels
        f = np.ctypeslib.as_array(self._L.mvit_slot_features(self._h, n, h, w, slot), shape=(n, c, h // 32, w // 32))
        p = np.ctypeslib.as_array(self._L.mvit_slot_pooled(self._h, n, h, w, slot), shape=(n, c))
        return f, p

    def profile(self, n, h, w, reps: int = 3) -> list:
        import json
        cap = 1 << 22
        buf = ctypes.create_string_buffer(cap)
        rc = self._L.mvit_profile_json(self._h, n, h, w, reps, buf, cap)
        if rc != 0:
            raise RuntimeError(f"mvit_profile_json rc={rc}")
        return json.loads(buf.value.decode())

    # ---- device-resident interface (bench.py) ----
    def prepare(self, n: int, h: int, w: int) -> None:
        if self._L.mvit_prepare(self._h, n, h, w) != 0:
            raise ValueError("mvit_prepare: invalid shape")

    def device_input(self, n, h, w) -> int:
        return self._L.mvit_device_input(self._h, n, h, w)

    def device_features(self, n, h, w) -> int:
        return self._L.mvit_device_features(self._h, n, h, w)

    def device_pooled(self, n, h, w) -> int:
        return self._L.mvit_device_pooled(self._h, n, h, w)

    def forward_device(self, n, h, w) -> None:
        if self._L.mvit_forward_device(self._h, n, h, w) != 0:
            raise ValueError("mvit_forward_device failed")

    def release(self, n, h, w) -> None:
        self._L.mvit_release(self._h, n, h, w)

    def plan_info(self, n, h, w) -> dict:
        info = PlanInfo()
        if self._L.mvit_plan_info(self._h, n, h, w, ctypes.byref(info)) != 0:
            raise ValueError("mvit_plan_info: invalid shape")
        return {k: getattr(info, k) for k, _ in PlanInfo._fields_}
