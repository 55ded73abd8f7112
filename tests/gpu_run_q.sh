#!/bin/bash
# full suite on the final code + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log; tail -4 gpurun_out/pytest_q.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_q_n1.json 2> gpurun_out/bench_q_n1.err; tail -c 400 gpurun_out/bench_q_n1.json; tail -3 gpurun_out/bench_q_n1.err
