"""Pretty-print parts of a bench.py JSON line: python tools/show_bench.py file.log [key ...]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
keys = sys.argv[2:] or [k for k in d if not isinstance(d[k], dict)]
head = {k: d.get(k) for k in ("metric", "value", "n_gpus", "ms_per_step", "moves_per_sec") if k in d}
print(json.dumps(head))
print("e2e", d.get("e2e", {}).get("value"))
for k in keys:
    print(k, json.dumps(d.get(k), indent=1))
