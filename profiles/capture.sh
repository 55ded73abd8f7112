#!/bin/bash
# Round profile capture, run on the B200 box:   gpurun --timeout 1500 -- 'bash profiles/capture.sh r1'
# Produces (in gpurun_out/, all small enough to be merged back):
#   launches_<tag>.csv      every kernel launch of `bench.py --steps 2 --warmup 3` with its device time (ncu, 1 metric)
#   metrics_<tag>.csv       ncu --set full raw page of the first 60 hot-kernel launches of one forward pass
#   top_<tag>.ncu-rep       full report incl. source of 3 launches of the dominant kernel
#   layers_<tag>.txt        per-launch table of the same plan (CUDA events), maps ncu launch i to bench.py's kernel names
# ncu runs only after the identical command exited 0 without it (B200_PROFILING.md).
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"k_gemm_tcgen05|k_dwconv|k_attention|k_layernorm|k_stem|k_vit_stage" -c 60 -o /tmp/step_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1 && \
ncu -i /tmp/step_$TAG.ncu-rep --page raw --csv > gpurun_out/metrics_$TAG.csv 2>/dev/null
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${2:-k_gemm_tcgen05}" -s 2 -c 3 -o gpurun_out/top_$TAG $CMD >> gpurun_out/ncu_full_$TAG.log 2>&1
python tests/profile_layers.py s 256 256 2>/dev/null | grep " us " > gpurun_out/layers_$TAG.txt
ls -la gpurun_out
