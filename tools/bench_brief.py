import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        b=json.loads(l); print(round(b["value"]), round(b["e2e"]["value"]), {k:round(v,2) for k,v in b["roofline"]["stage_ms_per_step"].items()}, [v["mismatching_frames"] for k,v in b["parity"].items() if isinstance(v,dict) and "mismatching_frames" in v], b.get("latency"), {k.split()[1] + k.split()[2]: round(v["frames_per_s"]) for k, v in (b.get("shapes") or {}).items()})
    else: print(l.rstrip()[:300])
