"""CPU: the arithmetic the kernels run (encoded size, CalculateWritableSize, GetFreeSize, GetWritableSize --
the B200_HD inlines of csrc/b200_dev.cuh, compiled here for the host) against the reference's own functions
(through oracle/_ref when built) and the plain-C oracle, on edge values and a random sweep."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import endpoint_lib
import orlib


@pytest.fixture(scope="module")
def dev():
    subprocess.check_call(["make", "-s", "-C", endpoint_lib.NATIVE, "libdev_helpers.so"])
    L = C.CDLL(os.path.join(endpoint_lib.NATIVE, "libdev_helpers.so"))
    u64 = C.c_uint64
    for name, n in (("dev_round_up8", 1), ("dev_encoded_size", 1), ("dev_calc_writable", 1), ("dev_free_size", 3),
                    ("dev_writable_size", 3), ("dev_sizeof_pairdev", 0), ("dev_sizeof_svccmd", 0)):
        f = getattr(L, name)
        f.restype = u64
        f.argtypes = [u64] * n
    return L


def test_layouts(dev):
    assert dev.dev_sizeof_pairdev() == 128 and dev.dev_sizeof_svccmd() == 128


def test_known_answers(dev):
    # SURVEY Appendix A, observed on the reference: CWS of {0,23,24,31,32,33,40,47,48} = {0,0,0,0,8,8,16,16,24}
    assert [dev.dev_calc_writable(s) for s in (0, 23, 24, 31, 32, 33, 40, 47, 48)] == [0, 0, 0, 0, 8, 8, 16, 16, 24]
    assert [dev.dev_encoded_size(p) for p in (0, 1, 7, 8, 9, 16, 16384, 16389)] == [16, 24, 24, 24, 32, 32, 16400, 16408]


def test_against_oracle_and_reference(dev, oracle):
    ref = orlib.Ref(debug=True) if orlib.ref_available(debug=True) else None
    rng = np.random.default_rng(5)
    vals = [0, 1, 7, 8, 9, 15, 16, 23, 24, 25, 31, 32, 33, 255, 256, 16384, 16389, (1 << 24) - 24, 1 << 24]
    vals += [int(x) for x in rng.integers(0, 1 << 25, 400)]
    for v in vals:
        assert dev.dev_round_up8(v) == oracle.L.orb_round_up(v)
        assert dev.dev_encoded_size(v) == oracle.L.orb_encoded_size(v)
        assert dev.dev_calc_writable(v) == oracle.L.orb_calc_writable(v)
        if ref is not None:
            if v > 0:   # the reference asserts payload_size > 0 (ring_buffer.h:181)
                assert dev.dev_encoded_size(v) == ref.L.ref_encoded_size(v)
            assert dev.dev_calc_writable(v) == ref.L.ref_calc_writable(v)
    ring_mem = np.zeros(1 << 16, dtype=np.uint8)
    ring = ref.L.ref_ring_create(ring_mem.ctypes.data, 1 << 16) if ref is not None else None
    for cap_log in (6, 10, 16, 24):
        cap = 1 << cap_log
        for _ in range(300):
            # the reference asserts an 8-byte aligned tail below the mask (ring_buffer.cc:100-101)
            h, t = int(rng.integers(0, cap)), int(rng.integers(0, cap // 8 - 1)) * 8
            assert dev.dev_free_size(cap, h, t) == oracle.L.orb_free_size(cap, h, t)
            assert dev.dev_writable_size(cap, h, t) == oracle.L.orb_writable_size(cap, h, t)
            if ref is not None and cap == 1 << 16:
                assert dev.dev_free_size(cap, h, t) == ref.L.ref_free_size(ring, h, t)
    if ring:
        ref.L.ref_ring_destroy(ring)


def test_service_wire_structs_and_eager_checksum():
    """Layout of what crosses PCIe between the host and the resident service warps, and the eager-push checksum:
    order-independent over lanes (the warp XOR-reduces lane-strided partial sums, the host walks the words),
    sensitive to every byte, to the size and to the delivered-count it was pushed for."""
    import ctypes as C
    import numpy as np
    subprocess.check_call(["make", "-s", "-C", endpoint_lib.NATIVE, "libdev_helpers.so"])
    D = C.CDLL(os.path.join(endpoint_lib.NATIVE, "libdev_helpers.so"))
    for f in ("dev_sizeof_svcdone", "dev_sizeof_eagerrec", "dev_offset_stamp2"):
        getattr(D, f).restype = C.c_uint64
    assert D.dev_sizeof_svcdone() == 16          # one 16-byte store: bytes, calls and stamp become visible together
    assert D.dev_sizeof_eagerrec() == 32
    assert D.dev_offset_stamp2() == 124          # second stamp in the second 64-byte half of the command line
    D.dev_eager_checksum.restype = C.c_uint64
    D.dev_eager_checksum.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_int]
    rng = np.random.default_rng(5)
    for size in (1, 7, 8, 9, 255, 1024, 2047, 2048):
        buf = rng.integers(0, 256, size + 8, dtype=np.uint8)
        ref = D.dev_eager_checksum(buf.ctypes.data, size, 12345, 1)
        assert D.dev_eager_checksum(buf.ctypes.data, size, 12345, 32) == ref      # warp form == host form
        assert D.dev_eager_checksum(buf.ctypes.data, size, 12346, 1) != ref       # another delivered count
        if size > 1:
            assert D.dev_eager_checksum(buf.ctypes.data, size - 1, 12345, 1) != ref
        b2 = buf.copy()
        b2[size - 1] ^= 1
        assert D.dev_eager_checksum(b2.ctypes.data, size, 12345, 1) != ref         # last byte counts
        b3 = buf.copy()
        b3[size] ^= 0xFF
        assert D.dev_eager_checksum(b3.ctypes.data, size, 12345, 1) == ref         # bytes past the frame do not
