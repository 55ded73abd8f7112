#!/bin/bash
mkdir -p gpurun_out
python tests/attn_one.py 256 32 144 > gpurun_out/b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_attention_tc -s 1 -c 1 -o gpurun_out/attn_tc python tests/attn_one.py 256 32 144 > gpurun_out/b_ncu.log 2>&1
echo "rc=$?"
