#!/bin/bash
# u8 stem route + vectorised f32 stem staging: targeted tests, then the default bench line and the per-launch table
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "u8 or stem or preprocess or split or batch_256" > gpurun_out/pytest_u8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_u8.log; tail -n 5 gpurun_out/pytest_u8.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_u8_n1.json 2> gpurun_out/bench_u8_n1.err; python - <<P
import json
d=json.loads(open('gpurun_out/bench_u8_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'f32', d['e2e']['f32_input']['value'], d['roofline']['frac'], d['clocks'])
print([ (k['kernel'],round(k['ms'],4)) for k in d.get('kernels',[])][:12])
P
python tests/profile_layers.py s 256 256 2>/dev/null | grep " us " | head -3
