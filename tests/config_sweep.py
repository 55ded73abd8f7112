"""Measures the other BASELINE.json configs (not the bench line): XS batch sweep 1-256 @256^2, S @512^2 batch 64, XXS batch 1.
Device-resident (CUDA-graph replay, wall clock over >= 20 forwards after warm-up) and parity vs the oracle on 2 images.
    python tests/config_sweep.py > gpurun_out/configs.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import weights as W
from oracle import binding

GF = {"s": 4.000, "xs": 2.051, "xxs": 0.814}

def run(variant, n, hw, reps):
    path = f"/tmp/w_{variant}.ggml"
    if not os.path.exists(path):
        W.write_weight_file(path, W.make_synthetic_weights(variant, 1234))
    m = G.MobileViT(path)
    imgs = W.synthetic_images(min(n, 8), hw, hw, seed=7)
    inp = m.host_input(n, hw, hw)
    for i in range(n):
        inp[i] = imgs[i % imgs.shape[0]]
    f, p = m.compute(n, hw, hw)
    k = min(n, 2)
    rf, rp = binding.OracleModel(path).forward(imgs[:k])
    rel = float(np.linalg.norm(f[:k] - rf) / np.linalg.norm(rf))
    top1 = bool((p[:k].argmax(1) == rp.argmax(1)).all())
    for _ in range(5):
        m.forward_device(n, hw, hw)
    G.lib_ggml().ggml_b200_synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.forward_device(n, hw, hw)
    G.lib_ggml().ggml_b200_synchronize()
    dt = (time.perf_counter() - t0) / reps
    info = m.plan_info(n, hw, hw)
    gf = GF[variant] * (hw / 256.0) ** 2
    print(f"MobileViT-{variant.upper():3s} {hw}x{hw} batch {n:4d}: {dt*1e3:8.3f} ms/forward {n/dt:10.0f} img/s  {n*gf/dt/1e3:7.1f} TFLOP/s  "
          f"mode={'fast' if info['mode']==0 else 'exact'} launches={info['launches']} arena={info['arena_bytes']/1e6:.0f}MB  relL2_vs_oracle={rel:.2e} top1={top1}", flush=True)
    m.close()

if __name__ == "__main__":
    run("xxs", 1, 256, 200)
    for b in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        run("xs", b, 256, 100 if b <= 32 else 30)
    run("s", 256, 256, 20)
    run("s", 64, 512, 20)
    run("s", 1, 256, 200)
