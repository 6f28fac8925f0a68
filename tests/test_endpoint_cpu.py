"""CPU: the product's endpoint state machine + BPEV hybrid poll loop (grpc-rdma_b200/host/
b200_endpoint.cc) driven by the reference's endpoint conformance pattern
(test/core/iomgr/endpoint_tests.cc) with the pair operations supplied by the CPU oracle through the
b200_pair_ops table -- host logic only, no GPU, no CUDA calls."""
import ctypes as C
import os
import re

import pytest

import endpoint_lib


@pytest.fixture(scope="module")
def drv(pkg):
    D, (O, ops) = endpoint_lib.load(pkg, need_oracle=True)
    return D, O, ops


def test_endpoint_library_exports_every_header_symbol(pkg):
    if not os.path.exists(pkg.ENDPOINT_LIB_PATH):
        pkg.build()
    txt = re.sub(r"/\*.*?\*/", "", open(pkg.ENDPOINT_HEADER).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(b200_(?:engine|endpoint|exchange)[a-z0-9_]*)\s*\(", txt)))
    assert len(syms) >= 14, syms
    C.CDLL(pkg.LIB_PATH, mode=C.RTLD_GLOBAL)
    L = C.CDLL(pkg.ENDPOINT_LIB_PATH)
    assert not [s for s in syms if not hasattr(L, s)]


@pytest.mark.parametrize("ring", [4096, 65536])
def test_read_and_write_8192_byte_slices(drv, ring):
    D, O, ops = drv
    O.oracle_ops_config(ring, 30)
    st = (C.c_uint64 * 4)()
    # endpoint_tests.cc:341 shape (10 MB / 100 kB writes / 8192-B slices), scaled to run in seconds
    assert D.drv_read_and_write(ops, 2_000_000, 100_000, 8192, 0, 50, 0, st) == 0
    assert st[0] > 0 and st[2] > 0  # the busy-poll scan synthesized the events


def test_read_and_write_one_byte_slices(drv):
    D, O, ops = drv
    O.oracle_ops_config(65536, 30)
    # :342 shape: every slice one byte -> one 24-byte ring frame per byte, <= max_sge per Send
    assert D.drv_read_and_write(ops, 60_000, 10_000, 1, 0, 50, 0, None) == 0


def test_read_and_write_with_shutdown(drv):
    D, O, ops = drv
    O.oracle_ops_config(65536, 30)
    assert D.drv_read_and_write(ops, 10_000_000, 100_000, 1, 1, 50, 0, None) == 0  # :343


def test_read_and_write_slice_size_sweep(drv):
    D, O, ops = drv
    O.oracle_ops_config(1024, 30)   # a ring smaller than one write: partial writes + credit returns
    i = 1
    while i < 1000:                 # :344-346
        assert D.drv_read_and_write(ops, 40320, i, i, 0, 50, 0, None) == 0, i
        i = max(i + 1, i * 5 // 4)


def test_message_larger_than_ring_and_staging(drv):
    D, O, ops = drv
    O.oracle_ops_config(4096, 30)
    assert D.drv_read_and_write(ops, 300_000, 300_000, 100_000, 0, 50, 0, None) == 0


def test_shutdown_sequence(drv):
    D, O, ops = drv
    O.oracle_ops_config(4096, 30)
    assert D.drv_shutdown_sequence(ops, 50) == 0


def test_peer_close_fails_pending_read(drv):
    D, O, ops = drv
    O.oracle_ops_config(4096, 30)
    assert D.drv_peer_close(ops, 50, 0) == 0


def test_echo_random_messages(drv):
    D, O, ops = drv
    O.oracle_ops_config(65536, 30)
    nbytes = C.c_uint64(0)
    # examples/cpp/test: random messages, msg == reply; one thread, one engine (oracle wire is not
    # thread-safe), sizes scaled to the small ring
    assert D.drv_echo(ops, 40, 300_000, 12345, 50, 0, 0, C.byref(nbytes)) == 0
    assert nbytes.value > 0


def test_many_connections_on_one_engine(drv):
    """120 connections in one pollable: more ready fds than one pass may synthesize events for
    (MAX_EPOLL_EVENTS = 100), requests of all connections in flight at once, echoes checked."""
    D, O, ops = drv
    O.oracle_ops_config(16384, 30)
    st = (C.c_uint64 * 4)()
    assert D.drv_multi_echo(ops, 120, 3, 20_000, 99, 50, 0, 0, st) == 0
    assert st[2] >= 120 * 3          # events synthesized by the busy-poll scan


def test_engine_lock_is_not_held_across_the_wait(drv):
    """One thread sleeps inside b200_engine_work(200 ms); another one calls b200_endpoint_write / read on the
    same engine: the calls return at once (the reference holds rdma_mu only around the pair scan,
    ev_epollex_rdma_bpev_linux.cc:1103-1145) and both complete."""
    D, O, ops = drv
    O.oracle_ops_config(65536, 30)
    worst = C.c_uint64(0)
    assert D.drv_two_threads(ops, 200, 5, C.byref(worst)) == 0
    assert worst.value < 50_000, "a call blocked for %d us behind the sleeping engine" % worst.value


@pytest.mark.parametrize("mode", ["batch", "async"])
def test_batching_engine_host_logic(drv, mode, monkeypatch):
    """The engine's batching paths on the CPU: an ops table WITH `submit` (one b200_pairs_submit-shaped call per
    pass: the rdma_flush / rdma_do_read loops of every queued endpoint) and one with the completion-queue form
    (post / poll, B200_ENDPOINT_ASYNC=1) over the oracle -- conformance shapes, a ring smaller than a write (partial
    writes re-armed through has_pending_writes), shutdown, peer close, echo and 120 connections on one engine."""
    D, O, _ = drv
    ops = O.oracle_pair_ops_batch() if mode == "batch" else O.oracle_pair_ops_async()
    if mode == "async":
        monkeypatch.setenv("B200_ENDPOINT_ASYNC", "1")
    O.oracle_ops_config(65536, 30)
    assert D.drv_read_and_write(ops, 2_000_000, 100_000, 8192, 0, 50, 0, None) == 0
    assert D.drv_read_and_write(ops, 60_000, 10_000, 1, 0, 50, 0, None) == 0
    assert D.drv_read_and_write(ops, 10_000_000, 100_000, 1, 1, 50, 0, None) == 0       # with shutdown
    O.oracle_ops_config(1024, 30)
    for i in (1, 7, 64, 513, 999):
        assert D.drv_read_and_write(ops, 40320, i, i, 0, 50, 0, None) == 0, i
    O.oracle_ops_config(4096, 30)
    assert D.drv_read_and_write(ops, 300_000, 300_000, 100_000, 0, 50, 0, None) == 0    # message > ring and staging
    assert D.drv_shutdown_sequence(ops, 50) == 0
    assert D.drv_peer_close(ops, 50, 0) == 0
    O.oracle_ops_config(65536, 30)
    nbytes = C.c_uint64(0)
    assert D.drv_echo(ops, 40, 300_000, 12345, 50, 0, 0, C.byref(nbytes)) == 0 and nbytes.value > 0
    O.oracle_ops_config(16384, 30)
    assert D.drv_multi_echo(ops, 120, 3, 20_000, 99, 50, 0, 0, None) == 0
