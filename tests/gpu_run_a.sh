#!/bin/bash
# full GPU test suite + bench A/B (IR fusion off / on)
mkdir -p gpurun_out; rm -f gpurun_out/a_summary.txt
timeout 2400 python -m pytest tests/ -q -m gpu > gpurun_out/a_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/a_summary.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_summary.txt
GGML_B200_IR_FUSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_irfuse.json 2> gpurun_out/a_bench_irfuse.err; echo "bench irfuse rc=$?" >> gpurun_out/a_summary.txt
cat gpurun_out/a_summary.txt
