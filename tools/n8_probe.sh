#!/bin/bash
# One 8-GPU gpurun call: host topology, then bench.py at N=8 with and without NUMA binding (copy bound inside the line).
mkdir -p gpurun_out
TAG=${1:-r2_n8}
{
  echo "== lscpu"; lscpu | grep -i -E "model name|socket|numa|^cpu\(s\)|thread|core"
  echo "== nodes"; ls /sys/devices/system/node/ 2>/dev/null
  for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist 2>/dev/null) mem $(grep MemTotal $n/meminfo 2>/dev/null | awk '{print $4}') kB"; done
  echo "== gpu numa"; for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do bb=$(echo $b | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "$b numa_node=$(cat /sys/bus/pci/devices/$bb/numa_node 2>/dev/null)"; done
  echo "== topo"; nvidia-smi topo -m
  echo "== affinity"; taskset -p $$; nproc; free -g | head -2
} > gpurun_out/${TAG}_topo.txt 2>&1
N=${2:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-latency --no-shapes"
$RUN > gpurun_out/${TAG}_bind.json 2> gpurun_out/${TAG}_bind.err; echo "bind rc=$?"
$RUN --no-bind --no-matcher > gpurun_out/${TAG}_nobind.json 2> gpurun_out/${TAG}_nobind.err; echo "nobind rc=$?"
tail -c 3000 gpurun_out/${TAG}_bind.json
