"""Tuning probe (not a test): times the tcgen05 GEMM on layer shapes.  python tests/gemm_probe.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ggml_experiments_b200 as G
L = G.lib_ggml()
L.ggml_b200_debug_gemm_time.restype = ctypes.c_float
L.ggml_b200_debug_gemm_time.argtypes = [ctypes.c_int] * 8
px = 256 * 128 * 128
shapes = [(px, 64, 16, 1, 1, 0, 0), (px, 32, 64, 0, 1, 0, 0), (px, 128, 32, 1, 1, 0, 0), (px // 4, 256, 64, 1, 1, 0, 0),
          (px // 4, 64, 256, 0, 1, 1, 1), (px // 16, 576, 144, 0, 1, 0, 0), (px // 16, 144, 144, 0, 0, 1, 1), (px // 16, 288, 144, 1, 1, 0, 0)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (M, N, K, act, w16, w32, res) in shapes:
    ms = L.ggml_b200_debug_gemm_time(M, N, K, act, w16, w32, res, 5)
    byt = M * K * 2 + M * N * (2 * w16 + 4 * w32 + 4 * res)
    print(f"M={M} N={N} K={K} act={act} f16={w16} f32={w32} res={res}: {ms*1e3:.1f} us  {byt/ms/1e6:.0f} GB/s", flush=True)
