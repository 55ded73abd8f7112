"""ctypes mirror of include/gru_b200.h: the batched GRU text generator (reference: rnn_text_gen/rnn_text_generation.cpp)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .mobilevit import lib_mobilevit

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


class GRU:
    def __init__(self, gru_bin_path: str):
        L = lib_mobilevit()
        L.gru_load.restype = ctypes.c_void_p
        L.gru_load.argtypes = [ctypes.c_char_p]
        L.gru_free.argtypes = [ctypes.c_void_p]
        L.gru_vocab.argtypes = [ctypes.c_void_p]
        L.gru_units.argtypes = [ctypes.c_void_p]
        L.gru_generate.restype = ctypes.c_float
        L.gru_generate.argtypes = [ctypes.c_void_p, _i32p, ctypes.c_int, ctypes.c_int, _i32p, _f32p]
        self._L = L
        self._h = L.gru_load(os.fsencode(gru_bin_path))
        if not self._h:
            raise FileNotFoundError(gru_bin_path)
        self.vocab, self.units = L.gru_vocab(self._h), L.gru_units(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gru_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def generate(self, first_tokens: np.ndarray, steps: int):
        """first_tokens [B] int32 -> (tokens [steps, B] int32, final_state [B, units] f32, ms for the whole loop)."""
        ft = np.ascontiguousarray(first_tokens, dtype=np.int32)
        B = ft.shape[0]
        out = np.empty((steps, B), dtype=np.int32)
        st = np.empty((B, self.units), dtype=np.float32)
        ms = self._L.gru_generate(self._h, ft.ctypes.data_as(_i32p), B, steps, out.ctypes.data_as(_i32p), st.ctypes.data_as(_f32p))
        if ms < 0:
            raise ValueError("gru_generate: invalid arguments")
        return out, st, float(ms)
