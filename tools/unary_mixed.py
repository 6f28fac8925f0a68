"""BASELINE configs[2] sweep and configs[4]-style mix on one GPU (documentation helper; profiles/):
  1. unary 1 KiB ping-pong through the service kernel for 1 .. 1024 connections (p50 / p99 / round trips per s);
  2. the same ping-pong on 128 connections WHILE 128 other connections stream 4 MiB messages through the batch
     kernels (k_send / k_recv launched beside the resident service kernel): both numbers under load.
No device-wide synchronisation anywhere: a persistent kernel is resident the whole time."""
import ctypes as C
import json
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
pkg.init(0)
L = pkg.lib()
PP = C.CDLL(os.path.join(os.path.dirname(pkg.LIB_PATH), "libb200_pingpong.so"))
PP.b200_pp_run.restype = C.c_double
PP.b200_pp_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]
workers = int(os.environ.get("SVC_WORKERS", "32"))
try:
    MAXG = max(1, len(os.sched_getaffinity(0)) // 2)
    q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
    if q != "max":
        MAXG = max(1, min(MAXG, int(int(q) / int(per)) // 2))
except Exception:
    MAXG = 8


def pct(r):
    r = np.sort(r.reshape(-1)) / 1e3
    return {"p50_us": round(float(r[len(r) // 2]), 2), "p99_us": round(float(r[int(len(r) * 0.99)]), 2)}


def pingpong(conns, groups, iters, m=1024):
    groups = min(groups, MAXG)          # client + server threads never exceed the cores of the box
    rtt = np.zeros(conns * iters, dtype=np.uint64)
    t = PP.b200_pp_run(conns, groups, iters, max(5, iters // 10), m, rtt.ctypes.data_as(C.POINTER(C.c_uint64)))
    d = pct(rtt) if t > 0 else {"error": int(t)}
    d.update({"connections": conns, "client_threads": groups, "server_threads": groups,
              "round_trips_per_s": round(conns * iters / t) if t > 0 else None})
    return d


dev = torch.device("cuda", 0)
# ---- streaming side (set up before the service starts: allocation-heavy)
sconns, msg = 128, 4 * 1024 * 1024
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 16384)
lens = pkg.chttp2_slice_lens(msg)
total = sum(lens)
pairs = [pkg.connected_pair("mx-tx%d" % c, "mx-rx%d" % c) for c in range(sconns)]
src = torch.randint(0, 255, (sconns * total,), dtype=torch.uint8, device=dev)
dst = torch.zeros(sconns * total, dtype=torch.uint8, device=dev)
sops, rops, keep = [], [], []
for c in range(sconns):
    off, sl = 0, []
    for n in lens:
        sl.append((src.data_ptr() + c * total + off, n))
        off += n
    arr = pkg.make_slices(sl)
    keep.append(arr)
    sops.append((pairs[c][0], arr, len(lens), 0))
    rops.append((pairs[c][1], dst.data_ptr() + c * total, total))
bs, br = pkg.Batch("send", sops, pkg.UNTIL_BLOCKED), pkg.Batch("recv", rops, pkg.UNTIL_BLOCKED)
stream = torch.cuda.Stream(device=dev)
sh = C.c_void_p(stream.cuda_stream)


def stream_steps(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        bs.launch(sh)
        br.launch(sh)
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / k


for _ in range(3):
    bs.launch(sh)
    br.launch(sh)
stream.synchronize()
alone_ms = stream_steps(20)

pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 4096)
assert L.b200_service_start(workers) == 0, pkg.last_error()
out = {"service_workers": workers, "unary_sweep": [], "mixed": None}
for conns, groups, iters in ((1, 1, 2000), (4, 2, 1000), (16, 4, 500), (64, 16, 200), (256, 32, 100), (1024, 32, 40)):
    out["unary_sweep"].append(pingpong(conns, groups, iters))
    print(json.dumps(out["unary_sweep"][-1]), flush=True)
res = {}
th = threading.Thread(target=lambda: res.update(pingpong(128, 16, 400)))
th.start()
ms = []
while th.is_alive():
    ms.append(stream_steps(10))
th.join()
ok = bs.results(sh) == [total] * sconns and br.results(sh) == [total] * sconns
stream.synchronize()
out["mixed"] = {"streaming": {"connections": sconns, "message_bytes": msg, "payload_GBps_alone": sconns * msg / alone_ms / 1e6,
                              "payload_GBps_beside_unary": sconns * msg / (sum(ms) / len(ms)) / 1e6, "intact": bool(ok)},
                "unary_beside_streaming": res}
print(json.dumps(out["mixed"]), flush=True)
L.b200_service_stop()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "unary_mixed.json"), "w"), indent=1)
