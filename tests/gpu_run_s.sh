#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "vit_" -x -q -s 2>&1 | grep -E "vit_mlp n=|passed|failed|rror" | head -20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s.log; tail -3 gpurun_out/pytest_s.log
for b in 32 256; do timeout 300 python tests/profile_layers.py s $b 256 > gpurun_out/layers_s_b${b}.txt 2>&1; done
head -1 gpurun_out/layers_s_*.txt; grep -E "vit_|linear" gpurun_out/layers_s_b256.txt | cut -c1-150
b() { timeout 300 python bench.py --batch $1 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; }
for bt in 1 32 256; do echo "== batch $bt"; b $bt; echo "== batch $bt MLP_FUSE=0"; GGML_B200_MLP_FUSE=0 b $bt; done
