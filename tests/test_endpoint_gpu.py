"""GPU: the endpoint + BPEV poll loop over the CUDA pair library (ops = NULL), driven by the
reference's conformance pattern (test/core/iomgr/endpoint_tests.cc) and its echo integration test
(examples/cpp/test).  Bytes travel host slice -> k_send -> HBM ring -> k_recv -> host slice."""
import ctypes as C

import pytest

import endpoint_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drv(gpu):
    D, _ = endpoint_lib.load(gpu, need_oracle=False)
    return D


def _ring(gpu, nbytes):
    gpu.config_set("B200_RING_BUFFER_SIZE_BYTES", nbytes)
    gpu.config_set("GRPC_RDMA_MAX_SGE", 30)


def test_read_and_write_8192_byte_slices(gpu, drv):
    _ring(gpu, 4 << 20)   # the reference default ring (config.cc:86-91)
    st = (C.c_uint64 * 4)()
    assert drv.drv_read_and_write(None, 10_000_000, 100_000, 8192, 0, 100, 0, st) == 0   # :341
    assert st[0] > 0


def test_read_and_write_one_byte_slices(gpu, drv):
    _ring(gpu, 65536)
    assert drv.drv_read_and_write(None, 20_000, 5_000, 1, 0, 100, 0, None) == 0          # :342 shape


def test_read_and_write_with_shutdown(gpu, drv):
    _ring(gpu, 65536)
    assert drv.drv_read_and_write(None, 100_000_000, 100_000, 1, 1, 100, 0, None) == 0   # :343


def test_read_and_write_slice_size_sweep(gpu, drv):
    _ring(gpu, 1024)
    i = 1
    while i < 1000:                                                                      # :344-346
        if i > 4:   # 1..4-byte slices make 10k-40k single-frame launches each; covered at 5+
            assert drv.drv_read_and_write(None, 40320, i, i, 0, 100, 0, None) == 0, i
        i = max(i + 1, i * 5 // 4)


def test_message_larger_than_ring_and_staging(gpu, drv):
    _ring(gpu, 65536)
    assert drv.drv_read_and_write(None, 3_000_000, 3_000_000, 100_000, 0, 100, 0, None) == 0


def test_shutdown_sequence(gpu, drv):
    _ring(gpu, 65536)
    assert drv.drv_shutdown_sequence(None, 100) == 0


@pytest.mark.parametrize("poller", [0, 1])
def test_peer_close_fails_pending_read(gpu, drv, poller):
    _ring(gpu, 65536)
    assert drv.drv_peer_close(None, 100, poller) == 0


def test_echo_client_and_server_threads_with_background_poller(gpu, drv):
    """examples/cpp/test: random messages of 1 B .. 4 MiB - 1 KiB (common.h:5-6), msg == reply;
    client and server on their own threads and engines, busy-poll window 0 so that readiness comes
    from the background poller's eventfd kicks through epoll_wait (the EV half of BPEV)."""
    _ring(gpu, 4 << 20)
    nbytes = C.c_uint64(0)
    assert drv.drv_echo(None, 24, 4 * 1024 * 1024 - 1024, 777, 0, 1, 1, C.byref(nbytes)) == 0
    assert nbytes.value > 0


def test_echo_busy_polling_only(gpu, drv):
    _ring(gpu, 1 << 20)
    assert drv.drv_echo(None, 24, 2_000_000, 4242, 200, 0, 1, None) == 0
