#!/bin/bash
# device throughput for VIORB_LANES x frames-per-pass
for l in 3 4; do for c in 96 128 192 256; do
  VIORB_LANES=$l python bench.py --steps 3 --warmup 3 --no-matcher --no-cpu --no-latency --chunk $c 2>/dev/null > /tmp/lane.json
  python - "$l" "$c" <<'PY'
import sys, json
d = json.load(open('/tmp/lane.json'))
print("lanes", sys.argv[1], "chunk", sys.argv[2], round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
done; done
