"""GPU: the persistent service kernel (b200_service_*).  While it runs, b200_pair_send / recv are
executed by resident worker CTAs from commands in pinned memory (no launch per call) and the
poller CTA keeps mirrors + ready ring current.  Same bit-exact bar as test_gpu_parity.py: every
single call's return value, the delivered bytes, all cursors and the ring image against the
golden fixtures and the oracle."""
import ctypes as C
import json
import os
import select
import time

import numpy as np
import pytest

import trace
from gpu_engine import GpuEngine

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "traces.json")))


@pytest.fixture(scope="module")
def svc(gpu):
    L = gpu.lib()
    assert L.b200_service_start(4) == 0, gpu.last_error()
    assert L.b200_service_running() == 4
    yield gpu
    L.b200_service_stop()
    assert L.b200_service_running() == 0


def _stats(L):
    out = (C.c_uint64 * 4)()
    L.b200_service_stats(out)
    return list(out)


def _compare(got, want, label):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, "%s: op %d (%s)\n got  %s\n want %s" % (label, i, w["op"], g, w)


@pytest.mark.parametrize("name", sorted(GOLDEN["traces"]))
def test_golden_traces_through_the_service(svc, name):
    t = GOLDEN["traces"][name]
    before = _stats(svc.lib())[0]
    recs = trace.run_trace(GpuEngine(svc, "pinned", 3), t["cap"], [tuple(o) for o in t["ops"]], GOLDEN["max_sge"])
    _compare(recs, t["records"], "golden %s [service]" % name)
    single = sum(1 for o in t["ops"] if o[0] in ("send", "recv"))
    if single:
        assert _stats(svc.lib())[0] > before  # the single calls really went through the resident kernel


def test_golden_full_size_through_the_service(svc):
    """The BASELINE-size fixture with the service running: the single Recv calls go through the owner warps /
    the pool, the rdma_flush / rdma_do_read loops through launches beside the resident kernels."""
    G = json.load(open(os.path.join(HERE, "golden", "traces_full.json")))
    for name, t in G["traces"].items():
        recs = trace.run_trace(GpuEngine(svc, "pinned", 3), t["cap"], [tuple(o) for o in t["ops"]], G["max_sge"],
                               ring_images=False)
        _compare(recs, t["records"], "golden full %s [service]" % name)


@pytest.mark.parametrize("seed", range(6))
def test_random_single_calls_vs_oracle(svc, oracle, seed):
    rng = np.random.default_rng(7000 + seed)
    cap = [64, 1024, 4096, 65536][seed % 4]
    ops = []
    for _ in range(60):
        if rng.integers(0, 2):
            n = int(rng.integers(1, 36))
            lens = [int(x) for x in rng.integers(0, min(3000, 2 * cap), n)]
            bidx = int(rng.integers(0, lens[0])) if lens[0] else 0
            ops.append(("send", lens, int(rng.integers(0, 1000)), bidx))
        else:
            ops.append(("recv", int(rng.integers(1, 2 * cap))))
    want = trace.run_trace(oracle, cap, ops)
    got = trace.run_trace(GpuEngine(svc, "device", seed), cap, ops)
    _compare(got, want, "service random seed %d cap %d" % (seed, cap))


def test_ping_pong_registered_memory(svc):
    """1 KiB unary ping-pong between two connected pairs with GPU-addressable host buffers
    (no bounce): the poller CTA's mirror refresh is what tells each side that data arrived."""
    pkg, L = svc, svc.lib()
    pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 64)
    a, b = pkg.connected_pair("pp-a", "pp-b")
    n = 1024
    buf = L.b200_mem_alloc_host(4 * n)
    tx = np.ctypeslib.as_array((C.c_uint8 * (4 * n)).from_address(buf))
    tx[:n] = np.arange(n, dtype=np.uint64).astype(np.uint8)
    sl_a = pkg.make_slices([(buf, n)])
    sl_b = pkg.make_slices([(buf + n, n)])
    rtts = []
    for it in range(200):
        tx[0] = it & 255
        t0 = time.perf_counter()
        assert L.b200_pair_send(a.h, sl_a, 1, 0) == n
        while not L.b200_pair_has_message(b.h):
            pass
        assert L.b200_pair_recv(b.h, buf + n, n) == n          # server receives into [n, 2n)
        assert L.b200_pair_send(b.h, sl_b, 1, 0) == n           # and echoes it
        while not L.b200_pair_has_message(a.h):
            pass
        assert L.b200_pair_recv(a.h, buf + 2 * n, n) == n
        rtts.append(time.perf_counter() - t0)
        assert np.array_equal(tx[2 * n:3 * n], tx[:n])
    rtts.sort()
    print("service ping-pong 1 KiB: p50 %.1f us, p99 %.1f us" % (rtts[100] * 1e6, rtts[197] * 1e6))
    a.disconnect()
    b.disconnect()
    L.b200_mem_free_host(buf)


def test_ready_ring_drives_the_background_poller(svc):
    """Poller::AddPollable with the service running: no scan launches; the device poller's ready
    ring entries become eventfd kicks (poller.cc:75-101), level-triggered like the reference."""
    pkg, L = svc, svc.lib()
    pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 64)
    a, b = pkg.connected_pair("rr-a", "rr-b")
    L.b200_poller_add(b.h)
    fd = b.wakeup_fd()
    launches = L.b200_launch_count()
    assert select.select([fd], [], [], 0.05)[0] == []           # idle: no kick
    payload = np.arange(300, dtype=np.uint64).astype(np.uint8)
    assert a.send([payload]) == 300
    assert select.select([fd], [], [], 5.0)[0] == [fd], "no eventfd kick from the ready ring"
    assert b.has_message() == 1 and b.readable() == 300
    L.b200_pair_consume_wakeup(b.h)
    assert select.select([fd], [], [], 2.0)[0] == [fd]          # still readable -> kicked again (level)
    got = b.recv(1000)
    assert np.array_equal(got, payload)
    L.b200_pair_consume_wakeup(b.h)
    time.sleep(0.05)
    L.b200_pair_consume_wakeup(b.h)
    assert select.select([fd], [], [], 0.1)[0] == []            # drained: quiet again
    assert L.b200_launch_count() == launches                    # nothing was launched for any of this
    st = _stats(L)
    assert st[1] >= 2 and st[2] == 0                            # entries consumed, no overrun
    # peer exit is a readiness change too (forces a read event, engine :1130-1137)
    a.disconnect()
    assert select.select([fd], [], [], 5.0)[0] == [fd]
    assert b.status() == 3
    L.b200_poller_remove(b.h)
    b.disconnect()


def test_endpoint_echo_with_the_service_running(svc):
    """The whole BPEV loop without a single launch: endpoint state machine + busy-poll/epoll engine
    (b200_endpoint.cc) over pairs whose Send / Recv are executed by the resident kernel and whose
    readiness comes from its poller CTA (eventfd kicks through the ready ring)."""
    import endpoint_lib
    pkg, L = svc, svc.lib()
    D, _ = endpoint_lib.load(pkg, need_oracle=False)
    pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 1024)
    launches = L.b200_launch_count()
    ops0 = _stats(L)[0]
    nbytes = C.c_uint64(0)
    # client and server threads, busy-poll window 100 us, background poller on
    assert D.drv_echo(None, 12, 1_500_000, 99, 100, 1, 1, C.byref(nbytes)) == 0
    assert nbytes.value > 0
    assert _stats(L)[0] > ops0 + 40             # the calls went through the service (batched: one op per pass and endpoint)
    assert L.b200_launch_count() == launches    # and nothing was launched
    # conformance shape: 100 kB writes of 8192-byte slices, byte ramp checked on the reader
    assert D.drv_read_and_write(None, 2_000_000, 100_000, 8192, 0, 100, 0, None) == 0
    assert L.b200_launch_count() == launches
