#!/bin/bash
# bench.py at N = 2 .. $1 GPUs of one box without the CPU / latency / shape blocks (scaling of `value`, `e2e` and the matcher only)
MAXN=${1:-4}; TAG=${2:-scaleq}
mkdir -p gpurun_out
for N in 2 4 8; do
  [ $N -le $MAXN ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620+N)) bench.py --gpus $N --steps 5 --warmup 3 \
    --no-cpu --no-latency --no-shapes > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err; echo "N=$N rc=$?"
done
