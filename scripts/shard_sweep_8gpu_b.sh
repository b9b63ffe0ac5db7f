# second 8-GPU pass: 36-qubit ENERGY (row-distributed sector matrix) and the 34-qubit energy+gradient with the final kernels
mkdir -p gpurun_out
OUT=gpurun_out/r02_shard_8gpu_b.jsonl
: > $OUT
run() { R=$1; shift; echo "== $R ranks: $*"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29585 scripts/shard_bench.py "$@" 2>&1 | tail -1 | tee -a $OUT; }
run 8 --n-orb 18 --what energy --steps 1 --warmup 0
run 8 --n-orb 17 --what grad --steps 1 --warmup 1
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
