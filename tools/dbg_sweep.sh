#!/bin/bash
# usage: tools/dbg_sweep.sh "0 2 3" [extra bench args]  -- prints stage times for VIORB_DEBUG masks
masks="$1"; shift
for d in $masks; do
  VIORB_DEBUG=$d python bench.py --steps 2 --warmup 3 --no-matcher --no-cpu --frames 1024 "$@" 2>/dev/null > /tmp/dbg_$d.json
  python - "$d" <<'PY'
import sys, json
d = json.load(open('/tmp/dbg_%s.json' % sys.argv[1]))
print("dbg", sys.argv[1], round(d["value"]), {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()})
PY
done
