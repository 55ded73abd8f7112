#!/bin/bash
# round-2 final evidence: whole GPU suite, smoke(), the default bench line, the per-GPU sub-batches of the strong-scaling points, the other
# BASELINE configs, the reference arm, ncu capture of the final code (compute-sanitizer is closed on this pool)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log; tail -n 3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 600 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 400 gpurun_out/bench_r2_n1.json; echo
for b in 128 64 32; do
  timeout 300 python bench.py --batch $b --no-cpu-baseline > gpurun_out/bench_r2_n1_b$b.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_r2_n1_b$b.json').read().strip().splitlines()[-1]); print($b, d['value'], d['ms_per_step'], d['e2e']['value'])"
done
for cfg in xs_sweep s512 gru; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_$cfg.json 2> gpurun_out/bench_r2_$cfg.err; tail -c 300 gpurun_out/bench_r2_$cfg.json; echo
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_r2_reference.json; echo
bash profiles/capture.sh r2 > gpurun_out/capture_final.log 2>&1
python tests/profile_layers.py s 32 256 2>/dev/null | grep " us \|total" > gpurun_out/layers_r2_b32.txt
ls -la gpurun_out | tail -30
