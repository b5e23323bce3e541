set -o pipefail
mkdir -p gpurun_out
T=${1:-d8}
timeout 900 python -m pytest tests/test_extract_gpu.py tests/test_cabi.py -x -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log; echo "pytest rc=$?"; cat gpurun_out/${T}_pytest.log
B="python bench.py --no-matcher --no-cpu --no-latency"
timeout 400 $B > gpurun_out/${T}_a.json 2> gpurun_out/${T}_a.err; echo "rc=$?"
cat gpurun_out/${T}_a.json | python tools/bench_brief.py
