TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581"
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sp2_list or large or 7-6-blocked" 2>&1 | tail -2
echo "== 1 GPU n=15 grad"; timeout 400 python scripts/shard_bench.py --group 1 --n-orb 15 --what grad --steps 1 --warmup 1 --check-norm 2>&1 | tail -1
echo "== 2 GPU n=15 grad"; timeout 400 $TR scripts/shard_bench.py --n-orb 15 --what grad --steps 1 --warmup 1 --check-norm 2>&1 | tail -1
echo "== 2 GPU n=15 grad nolist"; timeout 400 $TR scripts/shard_bench.py --n-orb 15 --what grad --steps 1 --warmup 1 --opt sp2_list=0 2>&1 | tail -1
echo "== 2 GPU n=15 dense prepare"; timeout 400 $TR scripts/shard_bench.py --n-orb 15 --what prepare --dense --steps 1 --warmup 1 2>&1 | tail -1
