#!/bin/bash
# fused transformer stage (vit_stage.cu): isolated parity, whole suite, per-launch tables with / without it, IR fusion at batch 32
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k vit_stage -x -q -s > gpurun_out/vit_l.log 2>&1; echo "vit rc=$?" | tee -a gpurun_out/vit_l.log
grep -E "vit_stage n=|passed|failed|Error|error" gpurun_out/vit_l.log | head -30
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_l.log; tail -5 gpurun_out/pytest_l.log
for b in 32 256; do
  timeout 300 python tests/profile_layers.py s $b 256 > gpurun_out/layers_l_b${b}.txt 2>&1
  GGML_B200_VIT_FUSE=0 timeout 300 python tests/profile_layers.py s $b 256 > gpurun_out/layers_l_b${b}_nofuse.txt 2>&1
  GGML_B200_VIT_FUSE=1 timeout 300 python tests/profile_layers.py s $b 256 > gpurun_out/layers_l_b${b}_fuse1.txt 2>&1
done
GGML_B200_IR_FUSE=1 timeout 300 python tests/profile_layers.py s 32 256 > gpurun_out/layers_l_b32_irfuse.txt 2>&1
head -1 gpurun_out/layers_l_*.txt
for b in 1 32 256; do
  echo "== bench batch $b"; timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
