b() { timeout 300 python bench.py --batch $1 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
for bt in 32 64; do
echo "== batch $bt"; b $bt
echo "== batch $bt lanes 2"; MVIT_LANES=2 b $bt
echo "== batch $bt lanes 4"; MVIT_LANES=4 b $bt
done
