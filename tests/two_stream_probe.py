"""Probe (not a test): device-resident throughput with two plans in flight on two streams vs one."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ggml_experiments_b200 as G
from ggml_experiments_b200 import weights as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
path = "/tmp/w_s.ggml"; W.write_weight_file(path, W.make_synthetic_weights("s", 1234))
m = G.MobileViT(path)
img = W.synthetic_images(1)[0]
for s in range(2):
    m.slot_input(n, 256, 256, s)[:] = img
    m.slot_submit(n, 256, 256, s); m.slot_wait(n, 256, 256, s)
    m.slot_set_transfers(n, 256, 256, s, False, False)
def run(k, slots):
    t0 = time.perf_counter()
    for i in range(k):
        s = i % slots
        m.slot_submit(n, 256, 256, s)
    for s in range(slots): m.slot_wait(n, 256, 256, s)
    return (time.perf_counter() - t0) / k
for slots in (1, 2):
    run(6, slots)
    dt = run(40, slots)
    print(f"batch {n} slots {slots}: {dt*1e3:.3f} ms/step {n/dt:.0f} img/s")
