#!/bin/bash
# round-2 re-entry: full GPU suite, bench line, per-launch tables at batch 256 / 32 / 512^2, then the ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k.log; tail -3 gpurun_out/pytest_k.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_k_n1.json 2> gpurun_out/bench_k_n1.err; tail -c 600 gpurun_out/bench_k_n1.json
python tests/profile_layers.py s 256 256 > gpurun_out/layers_k_b256.txt 2>&1
python tests/profile_layers.py s 32 256 > gpurun_out/layers_k_b32.txt 2>&1
python tests/profile_layers.py s 64 512 > gpurun_out/layers_k_s512.txt 2>&1
head -1 gpurun_out/layers_k_*.txt
bash profiles/capture.sh r2 > gpurun_out/capture_k.log 2>&1
ls gpurun_out | head -40
