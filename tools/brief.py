"""Print a one-line digest of a bench.py JSON line read from stdin (experiment helper)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not line:
    print(tag, "NO JSON")
    sys.exit(0)
d = json.loads(line[-1])
r = d.get("roofline") or {}
ks = {k: (round(v["ms"], 3), round(v["frac"], 3)) for k, v in (r.get("kernels") or {}).items()}
e = d.get("e2e") or {}
print(tag, "GB/s=%.1f ms/step=%.3f" % (d["value"], d["ms_per_step"]), ks,
      "e2e=%s %s" % (e.get("value") and round(e["value"], 1), {m: round(v["GBps"], 1) for m, v in (e.get("modes") or {}).items()}),
      "cpu=%s" % ((d.get("cpu_baseline") or {}).get("value")))
