"""Debug helper (not a test): per-stage comparison of the GPU path against the oracle.
    MVIT_DEBUG_STAGES=1 python tests/stage_debug.py [variant] [n] [hw] [fast|exact]"""
import os, sys
os.environ["MVIT_DEBUG_STAGES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import mobilevit as MV, weights as W
from oracle import binding

variant = sys.argv[1] if len(sys.argv) > 1 else "xxs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 64
mode = sys.argv[4] if len(sys.argv) > 4 else "exact"
path = f"/tmp/w_{variant}.ggml"
W.write_weight_file(path, W.make_synthetic_weights(variant, 1234))
imgs = W.synthetic_images(n, hw, hw, 7)
MV.set_mode(MV.FAST if mode == "fast" else MV.EXACT)
m = G.MobileViT(path)
feat, pooled = m.extract_features(imgs)
print("plan", m.plan_info(n, hw, hw))
om = binding.OracleModel(path)
names = ["stem", "layer1", "layer2", "layer3", "layer4", "layer5", "exp"]
for flags, tag in ((0, "oracle-f16"), (binding.PURE_F32, "oracle-f32")):
    print("---- vs", tag)
    for i in range(n):
        shapes = [m.debug_stage(n, hw, hw, s)[i].shape for s in range(7)]
        ref = om.forward_stages(imgs[i], shapes, flags)
        for s in range(7):
            got = m.debug_stage(n, hw, hw, s)[i]
            d = np.abs(got - ref[s])
            print(f"img {i} {names[s]:7s} shape {got.shape} relL2 {np.linalg.norm(got-ref[s])/np.linalg.norm(ref[s]):.3e} maxabs {d.max():.3e} refmax {np.abs(ref[s]).max():.3f}")
