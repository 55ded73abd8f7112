"""BASELINE.json config 5: GRU cell batched over 4096 independent streams, tokens/s on 1 B200 (not the bench.py line).
    python tests/gru_bench.py [B] [steps]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from ggml_experiments_b200.gru import GRU
from oracle import gru_oracle as GO
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = GO.make_synthetic_gru(seed=5)
path = os.path.join(tempfile.mkdtemp(), "gru.bin")
GO.write_gru_bin(path, w)
from ggml_experiments_b200 import mobilevit as MV
MV.set_mode(MV.EXACT if os.environ.get("GRU_MODE", "fast") == "exact" else MV.FAST)
m = GRU(path)
first = (np.arange(B) * 7 % 66).astype(np.int32)
m.generate(first, 3)  # warm-up (plan build)
toks, state, ms = m.generate(first, steps)
print('nan in final state:', int(np.isnan(state).sum()), 'token range', int(toks.min()), int(toks.max()), flush=True)
ref, margins, _ = GO.generate_batch(w, first[:64], min(steps, 20))
agree = float((toks[:min(steps, 20), :64] == ref).mean())
flop = 2.0 * B * (256 * 3072 + 1024 * 3072 + 1024 * 66)
print(f"GRU batched: B={B} steps={steps}: {ms:.2f} ms total, {ms/steps*1e3:.1f} us/step, {B*steps/ms*1e3:.0f} tokens/s, "
      f"{flop*steps/ms/1e9:.1f} TFLOP/s, first-20-step token agreement with the numpy oracle on 64 streams: {agree:.3f}")
