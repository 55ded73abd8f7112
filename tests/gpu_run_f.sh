#!/bin/bash
mkdir -p gpurun_out
GGML_B200_CUDA_GRAPH=0 python tests/fwd_once.py 32 256 2 > gpurun_out/f_plain.log 2>&1 &&
GGML_B200_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 90 --csv --log-file gpurun_out/f_launches_b32.csv python tests/fwd_once.py 32 256 2 > gpurun_out/f_ncu.log 2>&1
echo "rc=$?"
