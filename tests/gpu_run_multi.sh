#!/bin/bash
# strong-scaling bench line at N GPUs (BASELINE configs[2] as written: 256 images in total), launched the way the driver does
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n${N}_strong.json 2> gpurun_out/bench_r2_n${N}_strong.err
tail -c 300 gpurun_out/bench_r2_n${N}_strong.json; tail -3 gpurun_out/bench_r2_n${N}_strong.err
python - <<P
import json
d=json.loads(open('gpurun_out/bench_r2_n${N}_strong.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','scaling']}, d['e2e']['value'], d.get('gather'))
P
