#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -k conv3x3 -x -q 2>&1 | tail -2
GGML_B200_CONV_HALO=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k conv3x3 -x -q 2>&1 | tail -2
for b in 1 32 256; do
  echo "== bench batch $b"; timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
  echo "== bench batch $b IR_FUSE"; GGML_B200_IR_FUSE=1 timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
