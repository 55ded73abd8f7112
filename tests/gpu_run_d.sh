#!/bin/bash
mkdir -p gpurun_out
for b in 32 64 128; do for s in 1 2 4 8; do
  MVIT_LANES=$s timeout 300 python bench.py --steps 30 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/d_b${b}_s$s.json 2> gpurun_out/d_b${b}_s$s.err; echo "bench b$b lanes $s rc=$?"
done; done
timeout 2400 python -m pytest tests/ -q -m gpu -x > gpurun_out/d_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
