# launch list (durations) + full capture of the describe-stage kernels of a short run
set -o pipefail
mkdir -p gpurun_out
TAG=${1:-q}
KREGEX=${2:-blur_levels|describe_blurred}
CMD="python bench.py --steps 1 --warmup 1 --frames 256 --no-cpu --no-latency --no-shapes --no-matcher"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 2 -c 2 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu2 rc=$?"
