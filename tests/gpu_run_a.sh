#!/bin/bash
# round-2 GPU session A: isolated kernel tests (incl. the new fused inverted residual), tuning probe, whole-model parity, bench A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "fused_inverted" > gpurun_out/a_ir_tests.log 2>&1; echo "ir tests rc=$?" >> gpurun_out/a_summary.txt
GGML_B200_IR_KB64=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused_inverted" > gpurun_out/a_ir_tests_kb64.log 2>&1; echo "ir tests kb64 rc=$?" >> gpurun_out/a_summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not fused_inverted" > gpurun_out/a_kernel_tests.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/a_summary.txt
timeout 600 python tests/ir_probe.py 256 > gpurun_out/a_ir_probe.log 2>&1; echo "ir probe rc=$?" >> gpurun_out/a_summary.txt
timeout 1500 python -m pytest tests/ -q -m gpu --deselect tests/test_gpu_kernels.py > gpurun_out/a_gpu_tests.log 2>&1; echo "gpu tests (IR fuse on) rc=$?" >> gpurun_out/a_summary.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err; echo "bench fused rc=$?" >> gpurun_out/a_summary.txt
GGML_B200_NO_IR_FUSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_unfused.json 2> gpurun_out/a_bench_unfused.err; echo "bench unfused rc=$?" >> gpurun_out/a_summary.txt
cat gpurun_out/a_summary.txt
