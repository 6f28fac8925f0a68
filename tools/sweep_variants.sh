#!/bin/bash
# experiment helper: run the device-resident bench for every lib/libb200rdma_<name>.so variant
cd "$(dirname "$0")/.."
for lib in grpc-rdma_b200/lib/libb200rdma.so grpc-rdma_b200/lib/libb200rdma_*.so; do
  B200RDMA_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-unary "$@" 2>&1 | python -c '
import json,sys
name=sys.argv[1]
txt=sys.stdin.read()
try:
    d=json.loads(txt.strip().splitlines()[-1]); r=d["roofline"]["kernels"]
    print(name, "GBps %.1f" % d["value"], {k:(round(v["ms"],4), round(v["frac"],3)) for k,v in r.items()}, flush=True)
except Exception as e:
    print(name, "FAILED", txt[-400:])
' $(basename $lib)
done
