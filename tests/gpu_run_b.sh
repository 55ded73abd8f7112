#!/bin/bash
mkdir -p gpurun_out
python tests/ir_one.py L2b 8 16 256 2 > gpurun_out/b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ir_fused -s 1 -c 1 -o gpurun_out/ir_l2b_v4 python tests/ir_one.py L2b 8 16 256 2 > gpurun_out/b_ncu.log 2>&1
echo "rc=$?"
