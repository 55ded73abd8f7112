#!/bin/bash
# whole GPU suite, per-launch tables, default bench lines at 256 / 32 / 1 images
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_check.log; tail -n 3 gpurun_out/pytest_check.log
for b in 256 32; do python tests/profile_layers.py s $b 256 2>/dev/null > gpurun_out/layers_check_b$b.txt; head -1 gpurun_out/layers_check_b$b.txt; done
for b in 256 32 1; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; done
