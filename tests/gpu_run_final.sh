#!/bin/bash
# round-2 final evidence: sanitizer on the new kernels, the other BASELINE configs, ncu capture of the final code
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -k "vit_stage_fused_vs_float64 and (5-8-8 or 3-4-4 or 7-8-8)" -x -q > gpurun_out/sanitizer_vit_mem.log 2>&1; tail -4 gpurun_out/sanitizer_vit_mem.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -k "vit_stage_fused_vs_float64 and (5-8-8 or 3-16-16)" -x -q > gpurun_out/sanitizer_vit_race.log 2>&1; tail -4 gpurun_out/sanitizer_vit_race.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -k "conv3x3 and (1-8-8 or 3-16-16)" -x -q > gpurun_out/sanitizer_conv_mem.log 2>&1; tail -4 gpurun_out/sanitizer_conv_mem.log
for cfg in xs_sweep s512 gru; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_$cfg.json 2> gpurun_out/bench_r2_$cfg.err; tail -c 300 gpurun_out/bench_r2_$cfg.json; echo
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_r2_reference.json; echo
bash profiles/capture.sh r2 > gpurun_out/capture_final.log 2>&1
ls -la gpurun_out | tail -30
