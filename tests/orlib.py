"""ctypes bindings for the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

`Oracle`  -> oracle/liboracle.so        (plain-C restatement, rb_oracle.c)
`Ref`     -> oracle/_ref/libref_pair*.so (the reference's own ring_buffer.cc /
             pair.cc built unmodified over the loopback fake verbs)

Both expose the same small pair API so one trace can be replayed through
either.  Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class Slice(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_uint64)]


def make_slices(bufs):
    """bufs: list of numpy uint8 arrays (kept alive by the caller)."""
    arr = (Slice * max(1, len(bufs)))()
    for i, b in enumerate(bufs):
        arr[i].ptr = b.ctypes.data if b.size else 0
        arr[i].len = b.size
    return arr


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class OrbRing(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("capacity", C.c_uint64), ("mask", C.c_uint64),
                ("head", C.c_uint64), ("moving_head", C.c_uint64), ("remain", C.c_uint64)]


class OrbStatus(C.Structure):
    _fields_ = [("remote_head", C.c_uint64), ("peer_exit", C.c_int32), ("_pad", C.c_int32)]


class OrbPair(C.Structure):
    pass


OrbPair._fields_ = [
    ("ring", OrbRing), ("staging", C.c_void_p), ("staging_size", C.c_uint64),
    ("status_in", OrbStatus), ("status_out", OrbStatus),
    ("remote_tail", C.c_uint64), ("internal_read_size", C.c_uint64),
    ("partial_write", C.c_int), ("status", C.c_int), ("max_sge", C.c_int),
    ("peer", C.POINTER(OrbPair)), ("total_read", C.c_uint64), ("total_write", C.c_uint64),
    ("n_status_writes", C.c_uint64)]


class Oracle:
    """Plain-C port."""
    kind = "port"

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        u64 = C.c_uint64
        for name, res, args in [
            ("orb_round_up", u64, [u64]), ("orb_round_down", u64, [u64]),
            ("orb_encoded_size", u64, [u64]), ("orb_calc_writable", u64, [u64]),
            ("orb_free_size", u64, [u64, u64, u64]), ("orb_writable_size", u64, [u64, u64, u64]),
            ("orb_ring_init", None, [C.POINTER(OrbRing), C.c_void_p, u64]),
            ("orb_ring_has_message", C.c_int, [C.POINTER(OrbRing)]),
            ("orb_ring_readable", u64, [C.POINTER(OrbRing)]),
            ("orb_ring_read", u64, [C.POINTER(OrbRing), C.c_void_p, u64, C.POINTER(u64)]),
            ("orb_ring_place", u64, [C.c_void_p, u64, u64, C.c_void_p, u64]),
            ("orb_pair_create", C.POINTER(OrbPair), [u64, C.c_int]),
            ("orb_pair_destroy", None, [C.POINTER(OrbPair)]),
            ("orb_pair_connect", None, [C.POINTER(OrbPair), C.POINTER(OrbPair)]),
            ("orb_pair_send", u64, [C.POINTER(OrbPair), C.POINTER(Slice), C.c_size_t, C.c_size_t]),
            ("orb_pair_recv", u64, [C.POINTER(OrbPair), C.c_void_p, u64]),
            ("orb_pair_has_message", C.c_int, [C.POINTER(OrbPair)]),
            ("orb_pair_has_pending_writes", C.c_int, [C.POINTER(OrbPair)]),
            ("orb_pair_readable", u64, [C.POINTER(OrbPair)]),
            ("orb_pair_writable", u64, [C.POINTER(OrbPair)]),
            ("orb_pair_get_status", C.c_int, [C.POINTER(OrbPair)]),
            ("orb_pair_disconnect", None, [C.POINTER(OrbPair)]),
            ("orb_pair_send_all", u64, [C.POINTER(OrbPair), C.POINTER(Slice), C.c_size_t, C.c_size_t, C.POINTER(u64)]),
            ("orb_pair_recv_drain", u64, [C.POINTER(OrbPair), C.c_void_p, u64, C.POINTER(u64)]),
            ("orb_bench_stream", C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int, u64, C.POINTER(u64), C.c_size_t,
                                              C.POINTER(u64), C.POINTER(u64)]),
        ]:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args

    # uniform pair API ---------------------------------------------------
    def pair_pair(self, ring_capacity, max_sge=30):
        a = self.L.orb_pair_create(ring_capacity, max_sge)
        b = self.L.orb_pair_create(ring_capacity, max_sge)
        assert a and b
        self.L.orb_pair_connect(a, b)
        return a, b

    def destroy(self, p):
        self.L.orb_pair_destroy(p)

    def send(self, p, bufs, byte_idx=0):
        return self.L.orb_pair_send(p, make_slices(bufs), len(bufs), byte_idx)

    def send_all(self, p, bufs, byte_idx=0):
        calls = C.c_uint64(0)
        n = self.L.orb_pair_send_all(p, make_slices(bufs), len(bufs), byte_idx, C.byref(calls))
        return n, calls.value

    def recv(self, p, cap):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = self.L.orb_pair_recv(p, out.ctypes.data, cap)
        return out[:n].copy()

    def recv_drain(self, p, cap):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        calls = C.c_uint64(0)
        n = self.L.orb_pair_recv_drain(p, out.ctypes.data, cap, C.byref(calls))
        return out[:n].copy(), calls.value

    def state(self, p):
        s = p.contents
        return dict(head=s.ring.head, moving_head=s.ring.moving_head, remain=s.ring.remain,
                    remote_tail=s.remote_tail, internal_read_size=s.internal_read_size,
                    partial_write=int(s.partial_write), credit_remote_head=s.status_in.remote_head,
                    peer_exit=int(s.status_in.peer_exit))

    def ring_image(self, p):
        s = p.contents
        return np.ctypeslib.as_array((C.c_uint8 * s.ring.capacity).from_address(s.ring.buf)).copy()

    def has_message(self, p):
        return int(self.L.orb_pair_has_message(p))

    def has_pending_writes(self, p):
        return int(self.L.orb_pair_has_pending_writes(p))

    def readable(self, p):
        return self.L.orb_pair_readable(p)

    def writable(self, p):
        return self.L.orb_pair_writable(p)

    def status(self, p):
        return self.L.orb_pair_get_status(p)

    def disconnect(self, p):
        self.L.orb_pair_disconnect(p)

    def bench_stream(self, conns, threads, warm, msgs, ring_capacity, lens):
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        d, h = C.c_uint64(0), C.c_uint64(0)
        t = self.L.orb_bench_stream(conns, threads, warm, msgs, ring_capacity, _u64p(lens), lens.size,
                                    C.byref(d), C.byref(h))
        return t, d.value, h.value


def ref_available(debug=False):
    name = "libref_pair_dbg.so" if debug else "libref_pair.so"
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", name))


class Ref:
    """The reference's own PairPollable / RingBufferPollable (compiled unmodified)."""
    kind = "reference"

    def __init__(self, debug=True):
        name = "libref_pair_dbg.so" if debug else "libref_pair.so"
        L = self.L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", name))
        u64, vp = C.c_uint64, C.c_void_p
        for fname, res, args in [
            ("ref_set_ring_kb", None, [C.c_uint32]), ("ref_get_ring_kb", C.c_uint32, []),
            ("ref_pair_create", vp, []), ("ref_pair_destroy", None, [vp]),
            ("ref_pair_connect", C.c_int, [vp, vp]),
            ("ref_pair_address", C.c_size_t, [vp, vp, C.c_size_t]),
            ("ref_pair_send", u64, [vp, C.POINTER(Slice), C.c_size_t, C.c_size_t]),
            ("ref_pair_recv", u64, [vp, vp, u64]),
            ("ref_pair_has_message", C.c_int, [vp]), ("ref_pair_has_pending_writes", C.c_int, [vp]),
            ("ref_pair_readable", u64, [vp]), ("ref_pair_writable", u64, [vp]),
            ("ref_pair_get_status", C.c_int, [vp]), ("ref_pair_disconnect", None, [vp]),
            ("ref_pair_wakeup_fd", C.c_int, [vp]), ("ref_pair_max_sge", C.c_int, [vp]),
            ("ref_pair_state", None, [vp, C.POINTER(u64)]),
            ("ref_pair_ring", vp, [vp]), ("ref_pair_ring_size", u64, [vp]), ("ref_pair_staging", vp, [vp]),
            ("ref_pair_send_all", u64, [vp, C.POINTER(Slice), C.c_size_t, C.c_size_t, C.POINTER(u64)]),
            ("ref_pair_recv_drain", u64, [vp, vp, u64, C.POINTER(u64)]),
            ("ref_poller_add", None, [vp]), ("ref_poller_remove", None, [vp]),
            ("ref_ring_create", vp, [vp, u64]), ("ref_ring_destroy", None, [vp]),
            ("ref_ring_has_message", C.c_int, [vp]), ("ref_ring_readable", u64, [vp]),
            ("ref_ring_read", u64, [vp, vp, u64, C.POINTER(u64)]),
            ("ref_ring_state", None, [vp, C.POINTER(u64)]),
            ("ref_encoded_size", u64, [u64]), ("ref_calc_writable", u64, [u64]),
            ("ref_free_size", u64, [vp, u64, u64]),
            ("ref_ring_write_frames", u64, [vp, u64, vp, C.POINTER(Slice), C.c_size_t, C.POINTER(C.c_int)]),
            ("ref_bench_stream", C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int, u64, C.POINTER(u64), C.c_size_t,
                                              C.POINTER(u64), C.POINTER(u64)]),
            ("ref_bench_pingpong", C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int, u64, u64, C.POINTER(u64)]),
        ]:
            f = getattr(L, fname)
            f.restype = res
            f.argtypes = args

    def pair_pair(self, ring_capacity, max_sge=30):
        assert ring_capacity % 1024 == 0, "reference Config is in KB"
        assert max_sge == 30 or os.environ.get("FAKE_VERBS_MAX_SGE") == str(max_sge)
        self.L.ref_set_ring_kb(ring_capacity // 1024)
        a, b = self.L.ref_pair_create(), self.L.ref_pair_create()
        assert self.L.ref_pair_connect(a, b) == 1
        return a, b

    def destroy(self, p):
        self.L.ref_pair_disconnect(p)
        self.L.ref_pair_destroy(p)

    def send(self, p, bufs, byte_idx=0):
        return self.L.ref_pair_send(p, make_slices(bufs), len(bufs), byte_idx)

    def send_all(self, p, bufs, byte_idx=0):
        calls = C.c_uint64(0)
        n = self.L.ref_pair_send_all(p, make_slices(bufs), len(bufs), byte_idx, C.byref(calls))
        return n, calls.value

    def recv(self, p, cap):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = self.L.ref_pair_recv(p, out.ctypes.data, cap)
        return out[:n].copy()

    def recv_drain(self, p, cap):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        calls = C.c_uint64(0)
        n = self.L.ref_pair_recv_drain(p, out.ctypes.data, cap, C.byref(calls))
        return out[:n].copy(), calls.value

    def state(self, p):
        o = (C.c_uint64 * 8)()
        self.L.ref_pair_state(p, o)
        keys = ["head", "moving_head", "remain", "remote_tail", "internal_read_size", "partial_write",
                "credit_remote_head", "peer_exit"]
        return {k: int(v) for k, v in zip(keys, o)}

    def ring_image(self, p):
        n = self.L.ref_pair_ring_size(p)
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(self.L.ref_pair_ring(p))).copy()

    def has_message(self, p):
        return int(self.L.ref_pair_has_message(p))

    def has_pending_writes(self, p):
        return int(self.L.ref_pair_has_pending_writes(p))

    def readable(self, p):
        return self.L.ref_pair_readable(p)

    def writable(self, p):
        return self.L.ref_pair_writable(p)

    def status(self, p):
        return self.L.ref_pair_get_status(p)

    def disconnect(self, p):
        self.L.ref_pair_disconnect(p)

    def bench_pingpong(self, conns, groups, iters, warm, msg_bytes, ring_capacity):
        """Returns (wall seconds, rtt_ns[conns, iters]) of the reference's own Send/HasMessage/Recv ping-pong."""
        rtt = np.zeros(conns * iters, dtype=np.uint64)
        t = self.L.ref_bench_pingpong(conns, groups, iters, warm, msg_bytes, ring_capacity, _u64p(rtt))
        return t, rtt.reshape(conns, iters)

    def bench_stream(self, conns, threads, warm, msgs, ring_capacity, lens):
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        d, h = C.c_uint64(0), C.c_uint64(0)
        t = self.L.ref_bench_stream(conns, threads, warm, msgs, ring_capacity, _u64p(lens), lens.size,
                                    C.byref(d), C.byref(h))
        return t, d.value, h.value
