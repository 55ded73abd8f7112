"""Tuning probe (not a test): barrier-wait profile of the halo-mode conv3x3 (build with GGML_B200_GEMM_PROFILE=1).   python tests/conv_prof.py [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import weights as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
path = "/tmp/w_s.ggml"
W.write_weight_file(path, W.make_synthetic_weights("s", 1234))
m = G.MobileViT(path)
m.prepare(n, 256, 256)
m.host_input(n, 256, 256)[:] = W.synthetic_images(1, 256, 256)[0]
m.compute(n, 256, 256)
L = G.lib_ggml()
buf = (ctypes.c_ulonglong * 64)()
L.ggml_b200_debug_gemm_prof(buf, 16 + int(os.environ.get('NOLOAD', '0')))
m.profile(n, 256, 256, reps=1)  # warm-up pass + one timed pass = 2 launches of every kernel
L.ggml_b200_debug_gemm_prof(buf, 0)
names = ["32x32 C1=0", "32x32 fusion", "16x16 C1=0", "16x16 fusion"]
for s in range(4):
    g = [buf[16 * s + i] for i in range(8)]
    if not g[3]:
        continue
    c = g[3]
    print(f"{names[s]:13s} MMA role {g[7]/c:8.0f} clk (waits: tmem {g[4]/c:6.0f} A {g[5]/c:6.0f} B {g[6]/c:6.0f}) | producer {g[2]/c:8.0f} (waits: A-empty {g[0]/c:6.0f} B-empty {g[1]/c:6.0f})")
