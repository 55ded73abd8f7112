#!/bin/bash
# fused transformer stage: isolated parity + timing, whole suite, bench at batch 1 / 32 / 256
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k vit_stage -x -q -s 2>&1 | grep -E "vit_stage n=|passed|failed|rror" | head -30
python tests/vit_one.py 32 16 16 192 4 384 4 20 2>&1 | tail -1
python tests/vit_one.py 256 16 16 192 4 384 4 20 2>&1 | tail -1
python tests/vit_one.py 32 8 8 240 4 480 3 20 2>&1 | tail -1
python tests/vit_one.py 256 8 8 240 4 480 3 20 2>&1 | tail -1
if [ "$1" != "quick" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n.log; tail -3 gpurun_out/pytest_n.log
for b in 1 32 256; do
  echo "== bench batch $b"; timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
fi
