mkdir -p gpurun_out
OUT=gpurun_out/r02_shard_8gpu.jsonl
: > $OUT
run() { # ranks, args...
  R=$1; shift
  echo "== $R ranks: $*"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29582 scripts/shard_bench.py "$@" 2>&1 | tail -1 | tee -a $OUT
}
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
run 8 --n-orb 16 --what grad --steps 2 --warmup 1
run 8 --n-orb 17 --what grad --steps 1 --warmup 1
run 8 --n-orb 17 --what prepare --dense --steps 1 --warmup 1
run 8 --n-orb 18 --what prepare --steps 1 --warmup 1
run 8 --n-orb 18 --what prepare --dense --steps 1 --warmup 0
