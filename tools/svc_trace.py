"""Experiment helper: unary ping-pong on the -DB200_SVC_TRACE build, prints the owner warps' phase timers.
   B200RDMA_LIB=grpc-rdma_b200/lib/libb200rdma_trace.so python tools/svc_trace.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
L = pkg.lib()
pkg.init(0)
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 4096)
assert L.b200_service_start(4) == 0, pkg.last_error()
a, b = pkg.connected_pair("t-a", "t-b")
buf = L.b200_mem_alloc_host(8192)
n = 1024
sa = pkg.make_slices([(buf, n)])
sb = pkg.make_slices([(buf + 2048, n)])
import time
rt = []
for k in range(3000):
    t0 = time.perf_counter()
    assert L.b200_pair_send(a.h, sa, 1, 0) == n
    while not L.b200_pair_has_message(b.h):
        pass
    assert L.b200_pair_recv(b.h, buf + 2048, n) == n
    assert L.b200_pair_send(b.h, sb, 1, 0) == n
    while not L.b200_pair_has_message(a.h):
        pass
    assert L.b200_pair_recv(a.h, buf + 4096, n) == n
    rt.append(time.perf_counter() - t0)
rt.sort()
print("python-driven ping-pong p50 %.2f us" % (rt[len(rt) // 2] * 1e6))
out = (C.c_ulonglong * 16)()
L.b200_debug_service_trace.argtypes = [C.POINTER(C.c_ulonglong)]
if L.b200_debug_service_trace(out) == 0:
    t = list(out)
    ns, nans = max(1, t[2]), max(1, t[4])
    print("small sends %d: fetched->landed+pushed %.2f us; ->retire done %.2f us (of those with a retire); "
          "lines->ring stores %.2f us | all answered ops %d: fetched->answer %.2f us"
          % (t[2], t[0] / ns / 1e3, t[1] / ns / 1e3, t[5] / ns / 1e3, t[4], t[3] / nans / 1e3))
L.b200_service_stop()
