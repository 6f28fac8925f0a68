"""PCIe calibration: H2D alone, D2H alone, both at once (experiment helper; sets the e2e ceiling)."""
import subprocess
import time

import torch

n = 1 << 30
h_src = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dst = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_src, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_dst.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def chunks(k):
    step = n // k

    def f():
        for i in range(k):
            with torch.cuda.stream(s1):
                d_a[i * step:(i + 1) * step].copy_(h_src[i * step:(i + 1) * step], non_blocking=True)
            with torch.cuda.stream(s2):
                h_dst[i * step:(i + 1) * step].copy_(d_b[i * step:(i + 1) * step], non_blocking=True)
    return f


print("H2D alone   : %.1f GB/s" % (n / timed(h2d) / 1e9))
print("D2H alone   : %.1f GB/s" % (n / timed(d2h) / 1e9))
t = timed(both)
print("both at once: %.1f GB/s per direction (%.1f ms for 1 GiB each way)" % (n / t / 1e9, t * 1e3))
t = timed(chunks(256))
print("both, 256 x 4 MiB chunks: %.1f GB/s per direction" % (n / t / 1e9))
try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
    print(subprocess.run(["bash", "-c", "numactl -H | head -20; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv"],
                         capture_output=True, text=True).stdout)
except Exception as e:
    print(e)
