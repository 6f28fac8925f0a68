"""CPU: the product's endpoint state machine + BPEV poll loop (grpc-rdma_b200/host/b200_endpoint.cc) driving
the REFERENCE's own PairPollable and Poller (oracle/_ref/libref_pair_dbg.so: the reference's pair.cc /
ring_buffer.cc / poller.cc compiled unmodified, asserts ON, loopback fake verbs) through the b200_pair_ops
table.  The reference's invariants police every call the endpoint makes -- from two threads in the echo
tests, with the reference's Poller thread kicking the eventfds the engine sleeps on."""
import ctypes as C
import os
import subprocess

import pytest

import endpoint_lib

HERE = os.path.dirname(os.path.abspath(__file__))
REFLIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_pair_dbg.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def drv(pkg):
    D, _ = endpoint_lib.load(pkg, need_oracle=False)
    subprocess.check_call(["make", "-s", "-C", endpoint_lib.NATIVE, "libref_pair_ops.so"])
    R = C.CDLL(os.path.join(endpoint_lib.NATIVE, "libref_pair_ops.so"))
    R.ref_pair_ops.restype = C.c_void_p
    R.ref_ops_config.argtypes = [C.c_uint32]
    return D, R, R.ref_pair_ops()


def test_conformance_8192_byte_slices(drv):
    D, R, ops = drv
    R.ref_ops_config(64)
    st = (C.c_uint64 * 4)()
    assert D.drv_read_and_write(ops, 4_000_000, 100_000, 8192, 0, 50, 0, st) == 0   # endpoint_tests.cc:341 shape
    assert st[0] > 0


def test_conformance_one_byte_slices_and_sweep(drv):
    D, R, ops = drv
    R.ref_ops_config(64)
    assert D.drv_read_and_write(ops, 60_000, 10_000, 1, 0, 50, 0, None) == 0         # :342 shape
    R.ref_ops_config(1)                                                                 # 1 KiB ring: partial writes + credit
    i = 1
    while i < 1000:                                                                     # :344-346, every other size
        assert D.drv_read_and_write(ops, 40320, i, i, 0, 50, 0, None) == 0, i
        i = max(i + 1, i * 8 // 5)


def test_message_larger_than_ring(drv):
    D, R, ops = drv
    R.ref_ops_config(4)
    assert D.drv_read_and_write(ops, 300_000, 300_000, 100_000, 0, 50, 0, None) == 0


def test_shutdown_and_peer_close(drv):
    D, R, ops = drv
    R.ref_ops_config(4)
    assert D.drv_read_and_write(ops, 10_000_000, 100_000, 1, 1, 50, 0, None) == 0     # :343 (shutdown)
    assert D.drv_shutdown_sequence(ops, 50) == 0
    assert D.drv_peer_close(ops, 50, 0) == 0      # Disconnect -> peer_exit over the wire -> "Pair closed"
    assert D.drv_peer_close(ops, 50, 1) == 0      # same, noticed through the reference Poller's eventfd kick


@pytest.mark.parametrize("busy_us,poller", [(200, 0), (0, 1), (50, 1)])
def test_echo_two_threads(drv, busy_us, poller):
    """examples/cpp/test: random messages, msg == reply; client and server engines on their own threads.
    busy_us = 0 with the poller on: readiness comes only from epoll_wait on the pairs' eventfds."""
    D, R, ops = drv
    R.ref_ops_config(1024)
    nbytes = C.c_uint64(0)
    assert D.drv_echo(ops, 16, 3_000_000, 4711 + busy_us, busy_us, poller, 1, C.byref(nbytes)) == 0
    assert nbytes.value > 0


@pytest.mark.parametrize("busy_us,poller", [(100, 0), (0, 1)])
def test_many_connections_two_threads(drv, busy_us, poller):
    """10 connections per engine, client and server engines on their own threads, all requests in flight at
    once; with busy_us = 0 every wake-up comes from the reference Poller's eventfd kicks through epoll_wait."""
    D, R, ops = drv
    R.ref_ops_config(64)
    assert D.drv_multi_echo(ops, 10, 3, 150_000, 7 + busy_us, busy_us, poller, 1, None) == 0
