#!/bin/bash
# phase profile of the fused transformer stage (library built with GGML_B200_VIT_PROFILE=1) + its parity tests
timeout 600 python -m pytest tests/test_gpu_kernels.py -k vit_stage -x -q -s 2>&1 | grep -E "vit_stage n=|passed|failed|rror" | head -30
python tests/vit_one.py 32 16 16 192 4 384 4 1 2>&1 | tail -4
python tests/vit_one.py 32 8 8 240 4 480 3 1 2>&1 | tail -4
