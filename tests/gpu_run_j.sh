#!/bin/bash
# attention timing at the three S stages (batch 256) and the 512^2 first stage, tcgen05 kernel vs mma.sync kernel
for cfg in "256 32 144" "256 16 192" "256 8 240" "64 64 144"; do
  echo "== $cfg"; python tests/attn_one.py $cfg 2>&1 | tail -1
  GGML_B200_ATTN_NO_TC=1 python tests/attn_one.py $cfg 2>&1 | tail -1
done
