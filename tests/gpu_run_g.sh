#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err; echo "bench n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/g_bench_n4.json 2> gpurun_out/g_bench_n4.err; echo "bench n4 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --weak > gpurun_out/g_bench_n${N}_weak.json 2> gpurun_out/g_bench_n${N}_weak.err; echo "bench weak n$N rc=$?"
nproc; lscpu | grep "Model name"
