"""Request-trace replay shared by the golden generator, the CPU tests and the GPU parity tests.

A trace is (ring capacity, list of ops).  Ops:
    ("send", lens, seed, byte_idx)       one PairPollable::Send call
    ("send_all", lens, seed, byte_idx)   rdma_flush loop: Send until it accepts nothing
    ("stream", lens, seed, recv_cap)     closed loop of send_all / recv_drain until all delivered
    ("recv", cap)                        one PairPollable::Recv call
    ("recv_drain", cap)                  rdma_do_read loop: Recv until empty / dst full
Payload bytes are a closed-form function of (seed, slice index, byte index), so fixtures
do not depend on any RNG implementation.  Every op yields a record: return value(s), SHA-1
of the delivered bytes, both pairs' cursors, readiness answers and the SHA-1 of the
receiver's ring image with the frame pad bytes masked (pad bytes are whatever the
reference's staging buffer held; they are never delivered).
"""
import hashlib

import numpy as np


def gen_bytes(seed, k, n):
    i = np.arange(n, dtype=np.uint64)
    v = (i * np.uint64(197) + np.uint64(seed * 131 + k * 17) + (i >> np.uint64(7)) * np.uint64(31)) & np.uint64(255)
    return v.astype(np.uint8)


def make_bufs(lens, seed):
    return [gen_bytes(seed, k, int(n)) for k, n in enumerate(lens)]


def up8(v):
    return (v + 7) // 8 * 8


def mask_pads(img, st, cap):
    """Zero the pad bytes of every frame still in the ring image."""
    img = img.copy()

    def zero(a, b):  # [a, b) circular
        for pos in range(a, b):
            img[pos % cap] = 0

    def u64(pos):
        return int(img[pos:pos + 8].view(np.uint64)[0])

    if st["remain"] > 0:
        end = st["moving_head"] + st["remain"]
        zero(end, up8(end))
    pos = st["head"]
    for _ in range(cap // 24 + 2):
        hdr = u64(pos)
        if hdr == 0 or hdr > cap - 24:
            break
        end = pos + 8 + hdr
        zero(end, up8(end))
        foot = (pos + 8 + up8(hdr)) % cap
        if u64(foot) != 0xFFFFFFFFFFFFFFFF:
            break
        pos = (foot + 8) % cap
    return img


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_trace(eng, cap, ops, max_sge=30, ring_images=True):
    """eng: adapter with pair_pair/send/send_all/recv/recv_drain/state/... (tests/orlib.py, conftest)."""
    tx, rx = eng.pair_pair(cap, max_sge)
    recs = []
    try:
        for op in ops:
            rec = {"op": op[0]}
            if op[0] in ("send", "send_all"):
                _, lens, seed, bidx = op
                bufs = make_bufs(lens, seed)
                if op[0] == "send":
                    rec["ret"] = int(eng.send(tx, bufs, bidx))
                else:
                    r, calls = eng.send_all(tx, bufs, bidx)
                    rec["ret"], rec["calls"] = int(r), int(calls)
            elif op[0] == "recv":
                out = eng.recv(rx, op[1])
                rec["ret"], rec["sha"] = int(out.size), sha(out)
            elif op[0] == "recv_drain":
                out, calls = eng.recv_drain(rx, op[1])
                rec["ret"], rec["calls"], rec["sha"] = int(out.size), int(calls), sha(out)
            elif op[0] == "stream":
                # endpoint-style closed loop: flush as far as credit allows, drain, repeat
                _, lens, seed, rcap = op
                bufs = make_bufs(lens, seed)
                idx = bidx = rounds = 0
                parts = []
                total = sum(int(x) for x in lens)
                got = 0
                while got < total and rounds < 10000:
                    rounds += 1
                    if idx < len(bufs):
                        sent, _ = eng.send_all(tx, bufs[idx:], bidx)
                        while sent > 0:
                            left = bufs[idx].size - bidx
                            if sent >= left:
                                sent -= left
                                idx += 1
                                bidx = 0
                            else:
                                bidx += sent
                                sent = 0
                    out, _ = eng.recv_drain(rx, rcap)
                    got += out.size
                    parts.append(out)
                allb = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
                rec["ret"], rec["rounds"], rec["sha"] = int(allb.size), rounds, sha(allb)
                rec["intact"] = bool(np.array_equal(allb, np.concatenate(bufs)))
            else:
                raise ValueError(op)
            st_tx, st_rx = eng.state(tx), eng.state(rx)
            rec["tx"] = {k: st_tx[k] for k in ("remote_tail", "partial_write", "credit_remote_head")}
            rec["rx"] = {k: st_rx[k] for k in ("head", "moving_head", "remain", "internal_read_size")}
            rec["has_message"] = int(eng.has_message(rx))
            rec["pending"] = int(eng.has_pending_writes(tx))
            rec["readable"] = int(eng.readable(rx))
            rec["writable"] = int(eng.writable(tx))
            if ring_images:
                rec["ring"] = sha(mask_pads(eng.ring_image(rx), st_rx, cap))
            recs.append(rec)
    finally:
        eng.destroy(tx)
        eng.destroy(rx)
    return recs


# ---------------------------------------------------------------- the fixture set
# (the cases the reference's missing unit tests would pin)

def golden_traces():
    T = {}
    # frame images for characteristic payload sizes
    sizes = [1, 7, 8, 9, 15, 16, 255, 256, 16384, 16389]
    ops = []
    for i, p in enumerate(sizes):
        ops += [("send", [p], 100 + i, 0), ("recv", 1 << 20)]
    T["frame_sizes_64k"] = (65536, ops)
    # largest frames a ring / the staging buffer admit: C/2-24 fits, C-24 gets cut
    cap = 4096
    T["max_frames_4k"] = (cap, [("send", [cap // 2 - 24], 1, 0), ("recv_drain", cap),
                                ("send", [cap - 24], 2, 0), ("recv_drain", cap),
                                ("send_all", [cap - 24], 3, 0), ("recv_drain", cap),
                                ("send_all", [3 * cap], 4, 0), ("recv_drain", 4 * cap),
                                ("send_all", [3 * cap], 4, 2048), ("recv_drain", 4 * cap)])
    # wrap: walk the tail through every 8-byte offset in the last 64 bytes of a 1 KiB ring
    ops = []
    for k in range(40):
        ops += [("send", [9, 29 + 8 * (k % 5)], 200 + k, 0), ("recv_drain", 4096)]
    T["wrap_walk_1k"] = (1024, ops)
    # partial reads with small destinations
    ops = [("send", [100, 13, 64], 7, 0)]
    for c in [1, 7, 8, 91, 1, 12, 3, 61, 5]:
        ops.append(("recv", c))
    ops += [("recv", 100), ("send", [57], 8, 3), ("recv", 7), ("recv_drain", 1000)]
    T["partial_reads_1k"] = (1024, ops)
    # max_sge cut-off: 31+ slices need several Send calls
    T["max_sge_cut_64k"] = (65536, [("send", [11] * 45, 9, 0), ("recv_drain", 4096),
                                    ("send", [11] * 45, 9, 0), ("send", [11] * 15, 10, 4),
                                    ("recv_drain", 4096), ("send_all", [9, 300] * 40, 11, 0),
                                    ("recv_drain", 1 << 16)])
    # a zero-length slice in the middle stops the call (pair.cc:683-685)
    T["zero_len_slice_4k"] = (4096, [("send", [5, 6, 0, 7], 12, 0), ("recv_drain", 100),
                                     ("send", [0, 7], 13, 0), ("send_all", [8, 0, 8], 14, 0),
                                     ("recv_drain", 100)])
    # credit exhaustion and the C/2 credit return
    cap = 2048
    ops = []
    for k in range(12):
        ops += [("send_all", [9, 400], 300 + k, 0)]
    ops += [("recv", 50), ("send_all", [9, 400], 320, 0)]
    for k in range(8):
        ops += [("recv_drain", 700), ("send_all", [9, 400, 9, 77], 330 + k, 0)]
    ops += [("recv_drain", 1 << 16), ("send_all", [9, 400], 340, 0), ("recv_drain", 1 << 16)]
    T["credit_2k"] = (cap, ops)
    # chttp2-shaped message through a ring smaller than the message
    lens = []
    data = 5 + 300000
    while data > 0:
        n = min(16384, data)
        lens += [9, n]
        data -= n
    T["chttp2_300k_128k"] = (131072, [("stream", lens, 21, 50000), ("stream", lens, 22, 1 << 20),
                                      ("stream", [300005], 23, 4096)])
    return T


def golden_traces_full():
    """BASELINE-size fixture: 16 MiB ring, 4 MiB chttp2-shaped messages -- 514 slices each --
    enough of them that the ring wraps more than twice; the ring filled to the brim (credit exhaustion at full
    size), drained through single Recv calls and rdma_do_read loops of several sizes.  Replayed without ring
    images (hashing 16 MiB per op in Python is what would make this slow), every cursor and every delivered
    byte (SHA-1) per op."""
    lens = []
    data = 5 + 4 * 1024 * 1024
    while data > 0:
        n = min(16384, data)
        lens += [9, n]
        data -= n
    ops = [("send_all", lens, 400 + k, 0) for k in range(4)]        # the 4th no longer fits: partial write
    ops += [("recv", 50000), ("recv", 9), ("recv", 16384), ("recv_drain", 1 << 20), ("recv_drain", 3 << 20),
            ("send_all", lens, 404, 0),                            # credit came back at C/2: room again
            ("recv_drain", 1 << 25)]                               # everything
    for k, rcap in enumerate([1 << 20, 5 << 20, 300000, 1 << 22, 1 << 20, 7 << 20, 1 << 16]):
        ops.append(("stream", lens, 410 + k, rcap))                # closed loops: > 2 more laps of the ring
    return {"chttp2_4m_16m": (16 * 1024 * 1024, ops)}
