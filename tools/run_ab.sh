set -o pipefail
mkdir -p gpurun_out
T=${1:-d8}
timeout 900 python -m pytest tests/test_extract_gpu.py -x -q 2>&1 | tail -3
B="python bench.py --no-matcher --no-cpu --no-latency --no-shapes --steps 4"
timeout 400 $B > gpurun_out/${T}_x.json 2> gpurun_out/${T}_x.err; cat gpurun_out/${T}_x.json | python tools/bench_brief.py
