"""Calibration: what can a grid of `nctas` CTAs, each copying `per_cta` bytes with this library's movers, reach?
(experiment helper; the library variant is selected with B200RDMA_LIB, see tools/sweep_variants.sh)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
pkg.init(0)
L = pkg.lib()
total = 1 << 30
src = torch.randint(0, 255, (total + (1 << 20),), dtype=torch.uint8, device="cuda")
dst = torch.empty(total + (1 << 20), dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sh = C.c_void_p(stream.cuda_stream)


def run(nctas, per_cta, mis, item, mode, reps=5):
    stride = per_cta
    args = (dst.data_ptr(), src.data_ptr(), per_cta - 4096, stride, nctas, 0, mis, item, mode, sh)
    for _ in range(2):
        L.b200_probe_copy(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        L.b200_probe_copy(*args)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return (3 if mode & 2 else 2) * nctas * (per_cta - 4096) / ms / 1e6


shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]] or [(256, 4 << 20)]
for nctas, per_cta in shapes:
    for mis in (0, 8):
        for item, mode in ((4096, 0), (4096, 2), (2048, 0), (2048, 2)):
            print("%s ctas=%4d per_cta=%7d mis=%d item=%d %-10s %6.0f GB/s" % (
                os.path.basename(os.environ.get("B200RDMA_LIB", "default")), nctas, per_cta, mis, item,
                "copy+clear" if mode & 2 else "copy", run(nctas, per_cta, mis, item, mode)), flush=True)
