"""Debug helper (not a test): per-launch table of the fused plan.   python tests/profile_layers.py [variant] [batch] [hw]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import mobilevit as MV, weights as W
variant = sys.argv[1] if len(sys.argv) > 1 else "s"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 256
path = f"/tmp/w_{variant}.ggml"
W.write_weight_file(path, W.make_synthetic_weights(variant, 1234))
m = G.MobileViT(path)
m.prepare(n, hw, hw)
inp = m.host_input(n, hw, hw)
inp[:] = W.synthetic_images(1, hw, hw)[0]
m.compute(n, hw, hw)
prof = m.profile(n, hw, hw, reps=5)
tot = sum(r["ms"] for r in prof)
print(f"total {tot:.3f} ms over {len(prof)} launches")
for r in prof:
    gbs = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
    print(f"{r['ms']*1e3:8.1f} us {gbs:7.0f} GB/s {tf:7.1f} TF/s  {r['bytes']/1e6:8.1f} MB  {r['kernel']:32s} {r['what']}")
