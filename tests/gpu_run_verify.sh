#!/bin/bash
# final verification of the committed state: whole GPU suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_verify.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_verify.log; tail -3 gpurun_out/pytest_verify.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_verify_n1.json 2> gpurun_out/bench_verify_n1.err; python - <<P
import json
d=json.loads(open('gpurun_out/bench_verify_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','steps','warmup']}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
P
