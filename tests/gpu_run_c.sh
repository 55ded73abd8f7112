#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "fused_inverted" > gpurun_out/c_ir_tests.log 2>&1; echo "ir tests rc=$?"
for b in 32 8; do echo "== batch $b"; timeout 600 python tests/ir_probe.py $b 2>&1 | grep -v "^ir_fused\|unsupported"; done > gpurun_out/c_ir_probe_small.log 2>&1; echo "ir probe rc=$?"
