"""Generate tests/golden/traces.json from the REFERENCE's own code.

Runs every trace of tests/trace.golden_traces() through oracle/_ref/libref_pair_dbg.so
(the reference's ring_buffer.cc + pair.cc compiled unmodified, asserts ON, over the
loopback fake verbs) and stores the per-op records.  Needs /root/reference (to build
oracle/_ref); the committed JSON is what travels to the GPU box.

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orlib  # noqa: E402
import trace  # noqa: E402


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    R = orlib.Ref(debug=True)
    out = {"generator": "oracle/_ref/libref_pair_dbg.so (reference ring_buffer.cc + pair.cc, unmodified)",
           "max_sge": 30, "traces": {}}
    for name, (cap, ops) in trace.golden_traces().items():
        recs = trace.run_trace(R, cap, ops)
        out["traces"][name] = {"cap": cap, "ops": [list(o) for o in ops], "records": recs}
        print("%-22s cap=%-7d ops=%-3d" % (name, cap, len(ops)))
    path = os.path.join(HERE, "traces.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")
    # BASELINE-size traces: their own file, no ring images
    full = {"generator": out["generator"], "max_sge": 30, "ring_images": False, "traces": {}}
    for name, (cap, ops) in trace.golden_traces_full().items():
        recs = trace.run_trace(R, cap, ops, ring_images=False)
        full["traces"][name] = {"cap": cap, "ops": [list(o) for o in ops], "records": recs}
        print("%-22s cap=%-9d ops=%-3d" % (name, cap, len(ops)))
    path = os.path.join(HERE, "traces_full.json")
    with open(path, "w") as f:
        json.dump(full, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
