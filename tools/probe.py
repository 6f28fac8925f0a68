"""Calibration sweep: what can the one-CTA-per-connection copy decomposition reach? (experiment helper)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
pkg.init(0)
L = pkg.lib()
stride = int(os.environ.get('PROBE_STRIDE', 4 << 20))
nmax = 1184
src = torch.randint(0, 255, (nmax * stride + 4096,), dtype=torch.uint8, device="cuda")
dst = torch.empty(nmax * stride + 4096, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sh = C.c_void_p(stream.cuda_stream)


def run(nctas, threads, mis, item, dyn, per_cta=(4 << 20) - 4096, reps=5):
    for _ in range(2):
        L.b200_probe_copy(dst.data_ptr(), src.data_ptr(), per_cta, stride, nctas, threads, mis, item, dyn, sh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        L.b200_probe_copy(dst.data_ptr(), src.data_ptr(), per_cta, stride, nctas, threads, mis, item, dyn, sh)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return 2 * nctas * per_cta / ms / 1e6  # GB/s of traffic


a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(3):
    b.copy_(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5):
    b.copy_(a)
e1.record(stream)
torch.cuda.synchronize()
print("torch copy 1 GiB: %.0f GB/s" % (2 * (1 << 30) * 5 / e0.elapsed_time(e1) / 1e6))
del a, b
for nctas in (256, 296):
    for mis in (0, 8, 5):
        for item, mode in ((4096, 0), (4096, 2), (2048, 0)):
            g = run(nctas, 288, mis, item, mode)
            if mode & 2:
                g *= 1.5  # read + write + clear
            print("ctas=%4d mis=%d item=%5d (%s): %6.0f GB/s of traffic"
                  % (nctas, mis, item, "copy+clear" if mode & 2 else "copy", g), flush=True)
