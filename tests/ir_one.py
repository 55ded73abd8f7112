"""One configuration of the fused inverted-residual kernel (for ncu): python tests/ir_one.py BLOCK TH TW NT [reps]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggml_experiments_b200 as G

L = G.lib_ggml()
L.ggml_b200_debug_ir_time.restype = ctypes.c_float
L.ggml_b200_debug_ir_time.argtypes = [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_float)]
blocks = {"L1": (128, 16, 64, 32, 1, 0), "L2a": (128, 32, 128, 64, 2, 0), "L2b": (64, 64, 256, 64, 1, 1), "L3ds": (64, 64, 256, 96, 2, 0),
          "L4ds": (32, 96, 384, 128, 2, 0), "L5ds": (16, 128, 512, 160, 2, 0)}
name, th, tw, nt = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
B = int(os.environ.get("IR_BATCH", "256"))
os.environ["GGML_B200_IR_TH"], os.environ["GGML_B200_IR_TW"], os.environ["GGML_B200_IR_NT"] = th, tw, nt
hw, cin, e, cout, s, res = blocks[name]
ms = L.ggml_b200_debug_ir_time(B, hw, hw, cin, e, cout, s, res, reps, None)
print(f"{name} TH={th} TW={tw} NT={nt}: {1e3 * ms:.1f} us")
