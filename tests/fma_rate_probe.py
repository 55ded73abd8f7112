"""Tuning probe (not a test): warp-instructions per clock per SM of FFMA / FHFMA (fma.rn.f32.f16) / HFMA2.  python tests/fma_rate_probe.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ggml_experiments_b200 as G
L = G.lib_ggml()
L.ggml_b200_debug_fma_rate.restype = ctypes.c_float
L.ggml_b200_debug_fma_rate.argtypes = [ctypes.c_int, ctypes.c_int]
for mode, name in ((0, "FFMA"), (1, "FHFMA"), (2, "HFMA2")):
    print(f"{name:6s} {L.ggml_b200_debug_fma_rate(mode, 20000):.2f} warp-instr / clk / SM")
