#!/bin/bash
b() { timeout 300 python bench.py --batch $1 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; }
for bt in 1 32; do
echo "== batch $bt default"; b $bt
echo "== batch $bt NO_DEEP"; GGML_B200_GEMM_NO_DEEP=1 b $bt
echo "== batch $bt NO_PDL"; GGML_B200_NO_PDL=1 b $bt
echo "== batch $bt NO_HALO"; GGML_B200_CONV_NO_HALO=1 b $bt
echo "== batch $bt VIT_FUSE=0"; GGML_B200_VIT_FUSE=0 b $bt
done
