"""Pretty-prints a bench.py JSON line (file argument): headline keys, then one line per config."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "n_gpus")}, d.get("e2e"), (d.get("roofline") or {}).get("frac"))
for c in d.get("configs") or []:
    print(json.dumps(c)[:1000])
