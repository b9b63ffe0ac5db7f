for o in "" "--opt sp2_warps=2" "--opt grad_tile_bits=14" "--opt low_bits=0 --opt grad_tile_bits=13"; do
echo "== $o"; timeout 200 python bench.py --steps 10 --warmup 3 $o 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],2), d['roofline']['phase_ms'], d['config']['passes_per_sweep'], d['energy'])"; done
