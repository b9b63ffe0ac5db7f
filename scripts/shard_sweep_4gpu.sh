# 32-qubit energy + gradient on 1 / 2 / 4 ranks (strong scaling of the sharded state), one box
mkdir -p gpurun_out
OUT=gpurun_out/r02_shard_4gpu.jsonl
: > $OUT
run() { R=$1; shift; echo "== $R ranks: $*"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29583 scripts/shard_bench.py "$@" 2>&1 | tail -1 | tee -a $OUT; }
run 4 --n-orb 16 --what grad --steps 2 --warmup 1
run 2 --n-orb 16 --what grad --steps 2 --warmup 1
run 1 --n-orb 16 --what energy --steps 2 --warmup 1
run 4 --n-orb 16 --what energy --steps 2 --warmup 1
run 4 --n-orb 17 --what energy --steps 1 --warmup 1
