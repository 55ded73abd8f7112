#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" > gpurun_out/h_attn_tests.log 2>&1; echo "attn tests rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_tc.json 2> gpurun_out/h_bench_tc.err; echo "bench tc rc=$?"
timeout 300 python bench.py --config s512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_s512_tc.json 2> gpurun_out/h_s512_tc.err; echo "s512 tc rc=$?"
