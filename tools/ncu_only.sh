#!/bin/bash
# profiling half of tools/gpu_round.sh: launch list + full captures of a short, latency-free bench run
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-latency --no-shapes --map 2000000"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fast_|describe|blur_|pyr_|octree' -s 24 -c 13 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hamming|top2' -s 4 -c 4 -o gpurun_out/${TAG}_prof_match -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu2 rc=$?"
