"""Adapter that lets tests/trace.run_trace drive the CUDA path through the C ABI.

send / recv        -> b200_pair_send / b200_pair_recv with plain numpy memory
                      (exercises the pinned bounce staging of the single-call path)
send_all / drain   -> prepared batches (B200_BATCH_UNTIL_BLOCKED) on GPU-addressable
                      memory, deliberately misaligned by `misalign` bytes
"""
import ctypes as C

import numpy as np


class GpuEngine:
    kind = "cuda"

    def __init__(self, pkg, mem="device", misalign=0):
        self.pkg, self.L = pkg, pkg.lib()
        self.mem, self.mis = mem, misalign
        self.n = 0

    # ---- memory helpers
    def _alloc(self, nbytes):
        nbytes = nbytes + 64
        p = self.L.b200_mem_alloc_device(nbytes) if self.mem == "device" else self.L.b200_mem_alloc_host(nbytes)
        assert p, self.pkg.last_error()
        return p

    def _free(self, p):
        (self.L.b200_mem_free_device if self.mem == "device" else self.L.b200_mem_free_host)(p)

    def _upload(self, base, arr):
        if arr.size == 0:
            return
        if self.mem == "device":
            assert self.L.b200_memcpy(base, arr.ctypes.data, arr.size, 0, None) == 0
            assert self.L.b200_stream_sync(None) == 0
        else:
            C.memmove(base, arr.ctypes.data, arr.size)

    def _download(self, base, n):
        out = np.zeros(max(n, 1), dtype=np.uint8)
        if n:
            if self.mem == "device":
                assert self.L.b200_memcpy(out.ctypes.data, base, n, 1, None) == 0
                assert self.L.b200_stream_sync(None) == 0
            else:
                C.memmove(out.ctypes.data, base, n)
        return out[:n]

    # ---- engine API
    def pair_pair(self, cap, max_sge=30):
        self.pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", cap)
        self.pkg.config_set("GRPC_RDMA_MAX_SGE", max_sge)
        self.n += 1
        return self.pkg.connected_pair("tx%d" % self.n, "rx%d" % self.n)

    def destroy(self, p):
        p.disconnect()
        p.putback()

    def send(self, p, bufs, byte_idx=0):
        return p.send(bufs, byte_idx)

    def recv(self, p, cap):
        return p.recv(cap)

    def send_all(self, p, bufs, byte_idx=0):
        # pack slices back to back with an odd gap so consecutive slices get different alignments
        offs, off = [], self.mis
        for b in bufs:
            offs.append(off)
            off += b.size + 3
        base = self._alloc(off)
        flat = np.zeros(off + 1, dtype=np.uint8)
        for b, o in zip(bufs, offs):
            flat[o:o + b.size] = b
        self._upload(base, flat[:off])
        sl = self.pkg.make_slices([(base + o, b.size) for b, o in zip(bufs, offs)])
        bt = self.pkg.Batch("send", [(p, sl, len(bufs), byte_idx)], self.pkg.UNTIL_BLOCKED)
        bt.launch()
        res, calls = bt.results()[0], bt.calls()[0]
        bt.destroy()
        self._free(base)
        return res, calls

    def recv_drain(self, p, cap):
        base = self._alloc(cap + self.mis)
        bt = self.pkg.Batch("recv", [(p, base + self.mis, cap)], self.pkg.UNTIL_BLOCKED)
        bt.launch()
        n, calls = bt.results()[0], bt.calls()[0]
        bt.destroy()
        out = self._download(base + self.mis, n).copy()
        self._free(base)
        return out, calls

    def state(self, p):
        return p.state()

    def ring_image(self, p):
        return p.ring_image()

    def has_message(self, p):
        return p.has_message()

    def has_pending_writes(self, p):
        return p.has_pending_writes()

    def readable(self, p):
        return p.readable()

    def writable(self, p):
        return p.writable()

    def status(self, p):
        return p.status()
