#!/bin/bash
# warp-uniform role loops (+ conv3x3 pair mode): whole GPU suite, per-launch tables at 256 and 32 images
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_uniform.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_uniform.log; tail -n 3 gpurun_out/pytest_uniform.log
python tests/profile_layers.py s 256 256 2>/dev/null > gpurun_out/layers_uniform_b256.txt; head -1 gpurun_out/layers_uniform_b256.txt
GGML_B200_CONV_NO_PAIR=1 python tests/profile_layers.py s 256 256 2>/dev/null > gpurun_out/layers_uniform_nopair_b256.txt; head -1 gpurun_out/layers_uniform_nopair_b256.txt
python tests/profile_layers.py s 32 256 2>/dev/null > gpurun_out/layers_uniform_b32.txt; head -1 gpurun_out/layers_uniform_b32.txt
python tests/profile_layers.py s 1 256 2>/dev/null > gpurun_out/layers_uniform_b1.txt; head -1 gpurun_out/layers_uniform_b1.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
