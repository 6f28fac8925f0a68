"""GPU: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bit-exact bar: delivered bytes, Send/Recv return values, partial_write flags, every cursor
(head / moving_head / remain / remote_tail / internal_read_size / credit word), readiness
answers and the receiver's ring image (pad bytes masked) after every op.
"""
import ctypes as C
import json
import os
import select

import numpy as np
import pytest

import trace
from gpu_engine import GpuEngine

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "traces.json")))


def _ops(raw):
    return [tuple(o) for o in raw]


def _compare(got, want, label):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, "%s: op %d (%s)\n got  %s\n want %s" % (label, i, w["op"], g, w)


@pytest.mark.parametrize("name", sorted(GOLDEN["traces"]))
@pytest.mark.parametrize("mem,mis", [("device", 0), ("device", 5), ("pinned", 9)])
def test_golden_traces_on_gpu(gpu, name, mem, mis):
    t = GOLDEN["traces"][name]
    recs = trace.run_trace(GpuEngine(gpu, mem, mis), t["cap"], _ops(t["ops"]), GOLDEN["max_sge"])
    _compare(recs, t["records"], "golden %s [%s+%d]" % (name, mem, mis))


GOLDEN_FULL = json.load(open(os.path.join(HERE, "golden", "traces_full.json")))


@pytest.mark.parametrize("name", sorted(GOLDEN_FULL["traces"]))
@pytest.mark.parametrize("mem,mis", [("device", 0), ("pinned", 5)])
def test_golden_full_size_on_gpu(gpu, name, mem, mis):
    """The headline configuration against the reference itself: 16 MiB ring, 4 MiB chttp2-shaped messages
    (514 slices), ring filled to the brim and wrapped more than twice -- every return value, every cursor and
    the SHA-1 of everything delivered, op by op, as oracle/_ref produced them (tests/golden/make_golden.py)."""
    t = GOLDEN_FULL["traces"][name]
    recs = trace.run_trace(GpuEngine(gpu, mem, mis), t["cap"], _ops(t["ops"]), GOLDEN_FULL["max_sge"], ring_images=False)
    _compare(recs, t["records"], "golden full %s [%s+%d]" % (name, mem, mis))


def _random_ops(rng, cap, n_ops):
    ops = []
    for _ in range(n_ops):
        k = rng.integers(0, 5)
        if k < 2:
            n = int(rng.integers(1, 40))
            style = rng.integers(0, 4)
            if style == 0:
                lens = [int(x) for x in rng.integers(1, 64, n)]
            elif style == 1:
                lens = [9 if i % 2 == 0 else int(rng.integers(1, min(16385, cap))) for i in range(n)]
            elif style == 2:
                lens = [int(x) for x in rng.integers(1, 2 * cap, max(1, n // 8))]
            else:
                lens = [int(x) for x in rng.integers(0, 20, n)]
            bidx = int(rng.integers(0, lens[0])) if lens[0] else 0
            ops.append(("send" if k == 0 else "send_all", lens, int(rng.integers(0, 1000)), bidx))
        elif k == 2:
            ops.append(("recv", int(rng.integers(1, cap))))
        else:
            ops.append(("recv_drain", int(rng.integers(1, 2 * cap))))
    return ops


@pytest.mark.parametrize("seed", range(10))
def test_random_traces_vs_oracle(gpu, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    cap = [64, 1024, 2048, 4096, 65536][seed % 5]
    ops = _random_ops(rng, cap, 80)
    want = trace.run_trace(oracle, cap, ops)
    got = trace.run_trace(GpuEngine(gpu, "device", seed % 16), cap, ops)
    _compare(got, want, "random seed %d cap %d" % (seed, cap))


@pytest.mark.parametrize("max_sge", [1, 4, 32])
def test_other_max_sge(gpu, oracle, max_sge):
    ops = [("send", [7] * 50, 1, 0), ("send_all", [9, 100] * 30, 2, 3), ("recv_drain", 1 << 16),
           ("send_all", [9, 100] * 30, 3, 0), ("recv_drain", 1 << 16)]
    want = trace.run_trace(oracle, 16384, ops, max_sge)
    got = trace.run_trace(GpuEngine(gpu, "device", 1), 16384, ops, max_sge)
    _compare(got, want, "max_sge %d" % max_sge)


def test_every_relative_alignment(gpu, oracle):
    """All 16 source alignments x partial reads at all 16 destination alignments."""
    cap = 8192
    for mis in range(16):
        ops = [("send_all", [9, 1000 + mis, 37, 5], 50 + mis, mis % 9), ("recv", 3 + mis), ("recv_drain", 4000)]
        want = trace.run_trace(oracle, cap, ops)
        got = trace.run_trace(GpuEngine(gpu, "device", mis), cap, ops)
        _compare(got, want, "alignment %d" % mis)


def test_full_size_stream_properties(gpu):
    """BASELINE config 2 sizes (16 MiB ring, 4 MiB chttp2-shaped messages), several connections
    in one batch, enough messages to wrap the ring twice: round trip is the identity, returned
    counts add up, cursors stay consistent and the ring is all zero once drained."""
    pkg, L = gpu, gpu.lib()
    pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 16384)
    pkg.config_set("GRPC_RDMA_MAX_SGE", 30)
    nconn, msg = 4, 4 * 1024 * 1024
    lens = pkg.chttp2_slice_lens(msg)
    total = sum(lens)
    pairs = [pkg.connected_pair("fs-tx%d" % c, "fs-rx%d" % c) for c in range(nconn)]
    src = L.b200_mem_alloc_device(nconn * total)
    dst = L.b200_mem_alloc_device(nconn * total)
    host = np.zeros((nconn, total), dtype=np.uint8)
    for c in range(nconn):
        i = np.arange(total, dtype=np.uint64)
        host[c] = ((i * np.uint64(2654435761) >> np.uint64(13)) + np.uint64(131 * c)) & np.uint64(255)
    assert L.b200_memcpy(src, host.ctypes.data, host.size, 0, None) == 0
    L.b200_stream_sync(None)
    sops, rops, keep = [], [], []
    for c in range(nconn):
        off, sl = 0, []
        for n in lens:
            sl.append((src + c * total + off, n))
            off += n
        arr = pkg.make_slices(sl)
        keep.append(arr)
        sops.append((pairs[c][0], arr, len(lens), 0))
        rops.append((pairs[c][1], dst + c * total, total))
    bs = pkg.Batch("send", sops, pkg.UNTIL_BLOCKED)
    br = pkg.Batch("recv", rops, pkg.UNTIL_BLOCKED)
    enc = sum(16 + (n + 7) // 8 * 8 for n in lens)
    out = np.zeros_like(host)
    for it in range(9):  # 9 x ~4.01 MiB: wraps the 16 MiB ring twice
        assert L.b200_memcpy(dst, np.zeros(1, np.uint8).ctypes.data, 0, 0, None) == 0
        bs.launch()
        br.launch()
        assert bs.results() == [total] * nconn
        assert br.results() == [total] * nconn
        assert br.calls() == [len(lens)] * nconn
        assert L.b200_memcpy(out.ctypes.data, dst, out.size, 1, None) == 0
        L.b200_stream_sync(None)
        assert np.array_equal(out, host), "iteration %d" % it
        for tx, rx in pairs:
            st, sr = tx.state(), rx.state()
            assert st["remote_tail"] == (enc * (it + 1)) % (16 << 20)
            assert sr["head"] == sr["moving_head"] == st["remote_tail"] and sr["remain"] == 0
            assert st["partial_write"] == 0 and not rx.has_message()
    for _, rx in pairs:
        assert not rx.ring_image().any(), "ring must read as all zero after a full drain"
    bs.destroy()
    br.destroy()
    L.b200_mem_free_device(src)
    L.b200_mem_free_device(dst)
    for tx, rx in pairs:
        tx.disconnect(); rx.disconnect(); tx.putback(); rx.putback()


def test_message_larger_than_ring_and_staging(gpu, oracle):
    """16 MiB flat message through a 1 MiB ring: partial writes + credit returns (config 4's edge)."""
    cap = 1 << 20
    ops = [("stream", [3 * cap + 12345], 77, cap // 3), ("stream", [9, 16384] * 100, 78, 1 << 22)]
    want = trace.run_trace(oracle, cap, ops, ring_images=False)
    got = trace.run_trace(GpuEngine(gpu, "device", 3), cap, ops, ring_images=False)
    _compare(got, want, "message larger than ring")
    assert got[0]["intact"] and got[1]["intact"]


def test_lifecycle_and_errors(gpu):
    pkg = gpu
    pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 4096)
    a, b = pkg.Pair("lc-a"), pkg.Pair("lc-b")
    assert a.status() == 1 and len(a.address()) == 48
    assert pkg.lib().b200_pool_get(b"lc-a") == a.h
    # not connected yet: data path moves nothing (pair.cc:657, :266)
    assert a.send([np.ones(8, np.uint8)]) == 0 and a.recv(8).size == 0
    # wrong blob size / wrong tag / different ring size are refused (pair.cc:148-149)
    assert not a.connect(b.address()[:40])
    bad = bytearray(b.address()); bad[32] ^= 0xFF
    assert not a.connect(bytes(bad)) and "tag" in a.error()
    pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 8192)
    c = pkg.Pair("lc-c")
    assert not a.connect(c.address()) and "ring buffer size" in a.error()
    assert a.connect(b.address()) and b.connect(a.address())
    assert a.status() == 2 and b.status() == 2
    msg = np.arange(200, dtype=np.uint8)
    assert a.send([msg]) == 200
    assert b.has_message() and b.readable() == 200
    a.disconnect()
    assert a.status() == 4 and b.status() == 3          # peer sees HalfClosed
    assert np.array_equal(b.recv(1000), msg)            # what is already in the ring still drains
    b.disconnect()
    # Init re-arms a used pair (pair.cc:88-89)
    pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 4096)
    for p in (a, b):
        pkg.lib().b200_pair_init(p.h)
    assert a.status() == 1 and not a.ring_image().any()
    assert a.connect(b.address()) and b.connect(a.address())
    assert b.send([msg]) == 200 and np.array_equal(a.recv(1000), msg)
    for p in (a, b, c):
        p.disconnect(); p.putback()
    with pytest.raises(ValueError):
        pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 24)
    pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 3000)   # not a power of two (ring_buffer.cc:22)
    with pytest.raises(RuntimeError):
        pkg.Pair("lc-d")


def test_poller_scan_and_eventfd(gpu):
    pkg, L = gpu, gpu.lib()
    pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 1024)
    pairs = [pkg.connected_pair("ps-a%d" % i, "ps-b%d" % i) for i in range(40)]
    rx = [p[1] for p in pairs]
    tx = [p[0] for p in pairs]
    sent = set(range(0, 40, 3))
    for i in sent:
        assert tx[i].send([np.full(20, i, np.uint8)]) == 20
    # a sender with a partial write raises "writable" (poller.cc:84)
    big = np.zeros(5000, np.uint8)
    tx[1].send([big])
    assert tx[1].has_pending_writes()
    allp = rx + tx
    arr = (C.c_void_p * len(allp))(*[p.h for p in allp])
    ev = (C.c_uint32 * len(allp))()
    nready = L.b200_poller_scan(arr, len(allp), ev)
    want_rx = sent | {1}
    for i in range(40):
        assert bool(ev[i] & pkg.EV_READABLE) == (i in want_rx), i
    assert ev[40 + 1] & pkg.EV_WRITABLE
    assert nready == len(want_rx) + 1
    # background poller kicks the eventfd of ready pairs only (poller.cc:66-101)
    for p in rx:
        L.b200_poller_add(p.h)
    ready_fds = {rx[i].wakeup_fd() for i in want_rx}
    r, _, _ = select.select([p.wakeup_fd() for p in rx], [], [], 5.0)
    deadline = 50
    seen = set(r)
    while seen != ready_fds and deadline:
        r, _, _ = select.select([p.wakeup_fd() for p in rx], [], [], 0.1)
        seen |= set(r)
        deadline -= 1
    assert seen == ready_fds
    # half-closed pairs are reported readable so the endpoint can surface UNAVAILABLE
    tx[2].disconnect()
    nready = L.b200_poller_scan(arr, len(allp), ev)
    assert ev[2] & pkg.EV_READABLE
    L.b200_poller_shutdown()
    for p in rx:
        L.b200_poller_remove(p.h)
    for a, b in pairs:
        a.disconnect(); b.disconnect(); a.putback(); b.putback()


@pytest.mark.parametrize("extra", [0, 4])   # 0: host-staged lanes, 4: B200_BATCH_ZEROCOPY
def test_unprepared_batch_entry_points(gpu, oracle, extra):
    pkg, L = gpu, gpu.lib()
    pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 65536)
    n = 21
    pairs = [pkg.connected_pair("ub-a%d" % i, "ub-b%d" % i) for i in range(n)]
    lens = [9, 3000, 9, 17, 0, 20000]
    total = sum(lens)
    hsrc = L.b200_mem_alloc_host(n * total)
    hdst = L.b200_mem_alloc_host(n * total)
    src = np.ctypeslib.as_array((C.c_uint8 * (n * total)).from_address(hsrc))
    dst = np.ctypeslib.as_array((C.c_uint8 * (n * total)).from_address(hdst))
    src[:] = np.random.default_rng(5).integers(0, 256, n * total, dtype=np.uint8)
    dst[:] = 0
    sops = (pkg.SendOp * n)()
    rops = (pkg.RecvOp * n)()
    keep = []
    for i in range(n):
        off, sl = 0, []
        for ln in lens:
            sl.append((hsrc + i * total + off, ln))
            off += ln
        arr = pkg.make_slices(sl)
        keep.append(arr)
        sops[i].pair, sops[i].slices, sops[i].nslices, sops[i].byte_idx = pairs[i][0].h, arr, len(lens), 0
        rops[i].pair, rops[i].dst, rops[i].cap = pairs[i][1].h, hdst + i * total, total
    acc = (C.c_uint64 * n)()
    # the zero-length slice stops every Send call in front of it (pair.cc:683-685)
    first = sum(lens[:4])
    assert L.b200_pairs_send(sops, n, pkg.UNTIL_BLOCKED | extra, acc, None) == 0
    assert list(acc) == [first] * n
    assert L.b200_pairs_recv(rops, n, pkg.UNTIL_BLOCKED | extra, acc, None) == 0
    assert list(acc) == [first] * n
    for i in range(n):
        assert np.array_equal(src[i * total:i * total + first], dst[i * total:i * total + first])
    # a second round without the blocker, results come back in the caller's op order
    for i in range(n):
        sl = pkg.make_slices([(hsrc + i * total + first, 1000 + i)])
        keep.append(sl)
        sops[i].slices, sops[i].nslices = sl, 1
        rops[i].dst, rops[i].cap = hdst + i * total + first, 1000 + i
    assert L.b200_pairs_send(sops, n, pkg.UNTIL_BLOCKED | extra, acc, None) == 0
    assert list(acc) == [1000 + i for i in range(n)]
    assert L.b200_pairs_recv(rops, n, pkg.UNTIL_BLOCKED | extra, acc, None) == 0
    assert list(acc) == [1000 + i for i in range(n)]
    for i in range(n):
        a0 = i * total + first
        assert np.array_equal(src[a0:a0 + 1000 + i], dst[a0:a0 + 1000 + i])
    L.b200_mem_free_host(hsrc)
    L.b200_mem_free_host(hdst)
    for a, b in pairs:
        a.disconnect(); b.disconnect(); a.putback(); b.putback()
