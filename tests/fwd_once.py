"""A few device-resident forwards of MobileViT-S at a given batch (for ncu launch lists): python tests/fwd_once.py BATCH [HW] [REPS]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import weights as W
n = int(sys.argv[1]); hw = int(sys.argv[2]) if len(sys.argv) > 2 else 256; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variant = os.environ.get("MVIT_VARIANT", "s")
p = os.path.join(tempfile.mkdtemp(), "w.ggml")
W.write_weight_file(p, W.make_synthetic_weights(variant, seed=1234))
m = G.MobileViT(p)
m.host_input(n, hw, hw)[:] = W.synthetic_images(min(n, 8), hw, hw)[np.arange(n) % min(n, 8)]
m.compute(n, hw, hw)
for _ in range(reps):
    m.forward_device(n, hw, hw)
G.lib_ggml().ggml_b200_synchronize()
print("ok", m.plan_info(n, hw, hw))
