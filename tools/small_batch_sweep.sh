#!/bin/bash
# device / e2e throughput of one GPU for the per-rank batch sizes of the 1/2/4/8-GPU sweep (4096/N frames), by frames per pass
for f in ${FRAMES:-4096 1024 512 256}; do for c in ${CHUNKS:-32 48 64 96 128}; do
  python bench.py --steps 5 --warmup 3 --no-matcher --no-cpu --no-latency --frames $f --chunk $c 2>/dev/null > /tmp/sb.json
  python - "$f" "$c" <<'PY'
import sys, json
d = json.load(open('/tmp/sb.json'))
print("frames", sys.argv[1], "chunk", sys.argv[2], "device", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3))
PY
done; done
