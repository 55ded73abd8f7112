"""One attention launch at the bench's first-ViT-stage shape (for ncu): python tests/attn_one.py [N] [HW] [C]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ggml_experiments_b200 as G
L = G.lib_ggml()
u16p = ctypes.POINTER(ctypes.c_uint16)
L.ggml_b200_debug_attention.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u16p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
c = int(sys.argv[3]) if len(sys.argv) > 3 else 144
heads = 4
d = c // heads
dp = L.ggml_b200_debug_attention_dp(d)
rng = np.random.default_rng(0)
qkv = np.zeros((n, hw, hw, 3, heads, dp), np.float16)
qkv[..., :d] = rng.normal(size=(n, hw, hw, 3, heads, d)).astype(np.float16)
out = np.zeros((n, hw, hw, c), np.uint16)
for _ in range(3):
    assert L.ggml_b200_debug_attention(qkv.view(np.uint16).ctypes.data_as(u16p), n, hw, hw, c, heads, out.ctypes.data_as(u16p)) == 0
print("ok", float(np.abs(out.view(np.float16).astype(np.float32)).mean()))
L.ggml_b200_debug_attention_time.restype = ctypes.c_float
L.ggml_b200_debug_attention_time.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
print("us per launch", 1e3 * L.ggml_b200_debug_attention_time(qkv.view(np.uint16).ctypes.data_as(u16p), n, hw, hw, c, heads, 20))
