#!/bin/bash
# long differential fuzz of every entry point on a GPU box (tools/fuzz_*.py), logs under gpurun_out/
mkdir -p gpurun_out
T=${1:-fuzz}
timeout 1500 python tools/fuzz_extract.py ${2:-600} 31 > gpurun_out/${T}_extract.log 2>&1; echo "extract rc=$?"; tail -1 gpurun_out/${T}_extract.log
timeout 900 python tools/fuzz_match.py > gpurun_out/${T}_match.log 2>&1; echo "match rc=$?"; tail -1 gpurun_out/${T}_match.log
timeout 900 python tools/fuzz_search.py > gpurun_out/${T}_search.log 2>&1; echo "search rc=$?"; tail -1 gpurun_out/${T}_search.log
