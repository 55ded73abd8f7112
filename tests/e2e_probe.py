"""Tuning probe (not a test): pipelined e2e throughput vs number of slots in flight.   python tests/e2e_probe.py [batch] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import weights as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = w = 256
path = "/tmp/w_s.ggml"
W.write_weight_file(path, W.make_synthetic_weights("s", 1234))
m = G.MobileViT(path)
img = W.synthetic_images(8, h, w)
host = np.tile(img, (n // 8, 1, 1, 1))
for nslots in (1, 2, 3, 4):
    for s in range(nslots):
        m.slot_input(n, h, w, s)[:] = host
        m.slot_submit(n, h, w, s)
    for s in range(nslots):
        m.slot_wait(n, h, w, s)
    t0 = time.perf_counter()
    for i in range(steps):
        s = i % nslots
        if i >= nslots:
            m.slot_wait(n, h, w, s)
        m.slot_submit(n, h, w, s)
    for s in range(nslots):
        m.slot_wait(n, h, w, s)
    dt = time.perf_counter() - t0
    print(f"slots={nslots}: {dt/steps*1e3:.3f} ms/step  {n*steps/dt:.0f} img/s", flush=True)
