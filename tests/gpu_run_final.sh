#!/bin/bash
# round-2 final evidence: the other BASELINE configs, the reference arm, ncu capture of the final code (compute-sanitizer is closed on this pool)
mkdir -p gpurun_out
for cfg in xs_sweep s512 gru; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_$cfg.json 2> gpurun_out/bench_r2_$cfg.err; tail -c 300 gpurun_out/bench_r2_$cfg.json; echo
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_r2_reference.json; echo
bash profiles/capture.sh r2 > gpurun_out/capture_final.log 2>&1
ls -la gpurun_out | tail -30
