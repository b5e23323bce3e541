#!/bin/bash
# bench.py at N = 2 .. $1 GPUs of one box (one call of gpurun --gpus $1), one JSON line per N into gpurun_out/<tag>_n<N>.json
MAXN=${1:-4}; TAG=${2:-scale}
mkdir -p gpurun_out
for N in 2 4 8; do
  [ $N -le $MAXN ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err; echo "N=$N rc=$?"
done
