"""Debug helper (not a test): time / profile the fused transformer stage on one shape.  python tests/vit_one.py N H W C heads F layers [reps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_kernels import _vit_params, _p, f32p, u16p
n, h, w, c, heads, f, nl = [int(a) for a in sys.argv[1:8]]
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 20
lib = G.lib_ggml()
lib.ggml_b200_debug_vit_stage.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.POINTER(f32p), f32p, u16p, f32p, ctypes.c_int, f32p]
rng = np.random.default_rng(0)
x = rng.standard_normal((n, h, w, c)).astype(np.float32)
layers = _vit_params(rng, c, f, nl)
flat = [a for p in layers for a in p]
arr = (f32p * len(flat))(*[_p(a, f32p) for a in flat])
out32 = np.zeros((n, h, w, c), np.float32)
ms = np.zeros(1, np.float32)
rc = lib.ggml_b200_debug_vit_stage(_p(x, f32p), n, h, w, c, heads, f, nl, 1e-5, arr, _p(out32, f32p), None, None, reps, _p(ms, f32p))
tiles = (n * h * w + 127) // 128
print(f"vit_stage rc={rc} n={n} {h}x{w} C={c} F={f} layers={nl} tiles={tiles}: {ms[0]*1e3:.1f} us per launch, {ms[0]*1e3/nl:.1f} us per layer")
