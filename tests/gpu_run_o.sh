#!/bin/bash
# halo-mode 3x3 convolution: isolated parity, whole suite, per-launch tables, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -k conv3x3 -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_o.log; tail -3 gpurun_out/pytest_o.log
for b in 32 256; do
  timeout 300 python tests/profile_layers.py s $b 256 > gpurun_out/layers_o_b${b}.txt 2>&1
done
head -1 gpurun_out/layers_o_*.txt; grep conv3x3 gpurun_out/layers_o_*.txt | cut -c1-160
for b in 1 32 256; do
  echo "== bench batch $b"; timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
