#!/bin/bash
# One gpurun call: GPU tests, the default bench (both arms), then ncu launch list + full capture of a short run.
set -o pipefail
mkdir -p gpurun_out
TAG=${1:-r1_v6}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --frames 256 --no-cpu --no-latency --no-shapes --map 2000000"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fast_|describe|blur_|pyr_|octree' -s 24 -c 13 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hamming|top2' -s 4 -c 4 -o gpurun_out/${TAG}_prof_match -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu2 rc=$?"
cat gpurun_out/${TAG}_pytest.log
