#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or fused_inverted" > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"
GGML_B200_IR_FUSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_ir.json 2> gpurun_out/i_bench_ir.err; echo "bench ir rc=$?"
timeout 300 python bench.py --config s512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_s512.json 2> gpurun_out/i_s512.err; echo "s512 rc=$?"
timeout 600 python tests/ir_probe.py 256 2>&1 | grep -v "^ir_fused\|unsupported" > gpurun_out/i_ir_probe.log; echo "probe rc=$?"
