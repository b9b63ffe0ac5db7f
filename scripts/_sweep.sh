for o in "" "--opt sigma_stage=0"; do
echo "== $o"; timeout 200 python bench.py --steps 10 --warmup 3 $o 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],2), d['roofline']['phase_ms'], d['config']['passes_per_sweep'], d['energy'])"; done
