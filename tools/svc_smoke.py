"""Experiment helper: step-by-step smoke of the service (prints progress so a hang can be located)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge


def say(*a):
    print("[%.3f]" % time.time(), *a, flush=True)


pkg = ge.load_package()
L = pkg.lib()
pkg.init(0)
say("init ok")
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 64)
assert L.b200_service_start(4) == 0, pkg.last_error()
say("service started")
a, b = pkg.connected_pair("s-a", "s-b")
say("pairs connected")
x = (np.arange(1024) * 7 % 251).astype(np.uint8)
n = a.send([x])
say("send ->", n, "has_message", b.has_message(), "readable", b.readable())
y = b.recv(4096)
say("recv ->", y.size, "equal", bool(np.array_equal(x, y)), "eager hits", L.b200_service_eager_hits())
say("state", a.state(), b.state())
for k in range(5):
    x[0] = k
    assert a.send([x]) == 1024
    while not b.has_message():
        pass
    y = b.recv(4096)
    assert np.array_equal(x, y), k
say("5 more round ok; eager hits", L.b200_service_eager_hits())
big = (np.arange(300000) % 253).astype(np.uint8)
sent = a.send([big])
say("big send ->", sent)
got = []
while sum(g.size for g in got) < sent:
    g = b.recv(1 << 20)
    if g.size == 0:
        say("recv 0?!", b.has_message(), b.readable())
        break
    got.append(g)
say("big recv ->", sum(g.size for g in got), bool(np.array_equal(np.concatenate(got), big[:sent])))
L.b200_service_stop()
say("service stopped")
