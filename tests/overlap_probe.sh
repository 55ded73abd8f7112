#!/bin/bash
# do two forwards in flight (pipelined slots without the FIFO chain) beat back-to-back forwards?  e2e = 2 slots; value = 1 stream
b() { timeout 300 python bench.py --batch $1 --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e_u8', round(d['e2e']['value']), 'e2e_f32', round(d['e2e']['f32_input']['value']))"; }
for bt in 8 32 64 256; do
echo "== batch $bt FIFO"; b $bt
echo "== batch $bt OVERLAP"; GGML_B200_SLOT_OVERLAP=1 b $bt
done
