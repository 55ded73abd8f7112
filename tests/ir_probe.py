"""Tuning probe for the fused inverted-residual kernel (K4): time of the fused block vs the three kernels it replaces, for the
seven blocks of MobileViT-S at batch 256, over a few tile configurations (env GGML_B200_IR_TH / _TW / _NT).  GPU box only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggml_experiments_b200 as G

L = G.lib_ggml()
L.ggml_b200_debug_ir_time.restype = ctypes.c_float
L.ggml_b200_debug_ir_time.argtypes = [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_float)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
blocks = [("L1", 128, 16, 64, 32, 1, 0), ("L2a", 128, 32, 128, 64, 2, 0), ("L2b", 64, 64, 256, 64, 1, 1), ("L3ds", 64, 64, 256, 96, 2, 0),
          ("L4ds", 32, 96, 384, 128, 2, 0), ("L5ds", 16, 128, 512, 160, 2, 0)]
configs = {1: [(8, 16, 256), (16, 16, 512), (8, 32, 512), (16, 16, 256), (4, 32, 256), (8, 16, 512)],
           2: [(4, 8, 256), (8, 8, 512), (4, 16, 512), (8, 16, 512), (2, 16, 256), (4, 16, 256), (4, 8, 512)]}
for name, hw, cin, e, cout, s, res in blocks:
    unf = ctypes.c_float(0)
    first = True
    for th, tw, nt in configs[s]:
        os.environ["GGML_B200_IR_TH"], os.environ["GGML_B200_IR_TW"], os.environ["GGML_B200_IR_NT"] = str(th), str(tw), str(nt)
        ms = L.ggml_b200_debug_ir_time(B, hw, hw, cin, e, cout, s, res, 10, ctypes.byref(unf) if first else None)
        if first:
            print(f"{name}: unfused (expand + dw + reduce) {1e3 * unf.value:.1f} us", flush=True)
            first = False
        print(f"{name}: fused TH={th} TW={tw} NT={nt}: {1e3 * ms:.1f} us" if ms > 0 else f"{name}: fused TH={th} TW={tw} NT={nt}: unsupported", flush=True)
