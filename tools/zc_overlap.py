"""Experiment helper: ceiling of the SM-driven (zero-copy) PCIe path when both directions run at once.
k_send reads pinned host slices over PCIe on one stream while k_recv of the previous step writes pinned host
destinations on another stream (B200_BATCH_ZEROCOPY | CONCURRENT).  python tools/zc_overlap.py [conns] [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

conns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
pkg = ge.load_package()
L = pkg.lib()
pkg.init(0)
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 16384)
msg = 4 << 20
lens = pkg.chttp2_slice_lens(msg)
total = sum(lens)
pairs = [pkg.connected_pair("z%d-tx" % c, "z%d-rx" % c) for c in range(conns)]
dstride = (total + 255) // 256 * 256
hsrc = L.b200_mem_alloc_host(conns * total)
hdst = L.b200_mem_alloc_host(conns * dstride)
hs = np.ctypeslib.as_array((C.c_uint8 * (conns * total)).from_address(hsrc))
hs[:] = np.arange(conns * total, dtype=np.uint64).astype(np.uint8)
FL = pkg.UNTIL_BLOCKED | pkg.ZEROCOPY | 0x8
sops, rops, keep = [], [], []
for c in range(conns):
    off, sl = 0, []
    for n in lens:
        sl.append((hsrc + c * total + off, n))
        off += n
    arr = pkg.make_slices(sl)
    keep.append(arr)
    sops.append((pairs[c][0], arr, len(lens), 0))
    rops.append((pairs[c][1], hdst + c * dstride, total))
bs, br = pkg.Batch("send", sops, FL), pkg.Batch("recv", rops, FL)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h1, h2 = C.c_void_p(s1.cuda_stream), C.c_void_p(s2.cuda_stream)
sent = [torch.cuda.Event() for _ in range(steps + 3)]
recvd = [torch.cuda.Event() for _ in range(steps + 3)]


def run(n):
    for k in range(n):
        if k >= 2:
            s1.wait_event(recvd[k - 2])      # credit: at most two messages outstanding per ring
        bs.launch(h1)
        sent[k].record(s1)
        s2.wait_event(sent[k])
        br.launch(h2)
        recvd[k].record(s2)
    torch.cuda.synchronize()


run(3)
t0 = time.perf_counter()
run(steps)
t = time.perf_counter() - t0
ok = bool(np.array_equal(hs.reshape(conns, total), np.ctypeslib.as_array((C.c_uint8 * (conns * dstride)).from_address(hdst)).reshape(conns, dstride)[:, :total]))
print("zero-copy, both directions overlapped: %.1f GB/s payload (%d conns, %d steps), intact=%s" % (conns * steps * msg / t / 1e9, conns, steps, ok))
# one direction at a time for comparison
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(steps):
    bs.launch(h1)
    torch.cuda.synchronize()
    br.launch(h1)
    torch.cuda.synchronize()
t = time.perf_counter() - t0
print("zero-copy, serialized: %.1f GB/s payload" % (conns * steps * msg / t / 1e9))
