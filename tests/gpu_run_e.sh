#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -m gpu -k "two_rank or lanes" > gpurun_out/e_multi_tests.log 2>&1; echo "multi tests rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench_n$N.json 2> gpurun_out/e_bench_n$N.err; echo "bench n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --weak > gpurun_out/e_bench_n${N}_weak.json 2> gpurun_out/e_bench_n${N}_weak.err; echo "bench weak n$N rc=$?"
