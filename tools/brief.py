"""Dev tool: print the key numbers of a bench.py JSON line read from stdin."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, d["value"], d["ms_per_step"], d["stage_ms"], "e2e", d["e2e"]["value"])
