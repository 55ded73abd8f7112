#!/bin/bash
# last check of the committed state: whole GPU suite, default bench line, S at 512x512
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_last.log; tail -n 3 gpurun_out/pytest_last.log
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/bench_last_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_last_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
timeout 40 python bench.py --config s512 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last_s512.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_last_s512.json').read().strip().splitlines()[-1]); print('s512', d['value'], d['ms_per_step'])"
