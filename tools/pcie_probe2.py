"""PCIe calibration 2: does the shape of the e2e path's copies (256 per direction, 4,196,622 bytes each, at
offsets that are not even 16-byte aligned, spread over 8+8 streams) cost bandwidth? (experiment helper)"""
import time

import torch

N, SZ = 256, 4196622
h_src = torch.empty(N * SZ + 4096, dtype=torch.uint8).pin_memory()
h_dst = torch.empty(N * SZ + 4096, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N * (SZ + 512) + 4096, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N * (SZ + 512) + 4096, dtype=torch.uint8, device="cuda")
ups = [torch.cuda.Stream() for _ in range(8)]
downs = [torch.cuda.Stream() for _ in range(8)]


def run(aligned, nstreams, reps=4, both=True):
    stride_h = ((SZ + 4095) // 4096) * 4096 if aligned else SZ
    if aligned:
        n = (h_src.numel() - 4096) // stride_h
    else:
        n = N

    def once():
        for c in range(n):
            ho = c * stride_h
            do = c * (SZ + 512) if not aligned else c * stride_h
            with torch.cuda.stream(ups[c % nstreams]):
                d_a[do:do + SZ].copy_(h_src[ho:ho + SZ], non_blocking=True)
            if both:
                with torch.cuda.stream(downs[c % nstreams]):
                    h_dst[ho:ho + SZ].copy_(d_b[do:do + SZ], non_blocking=True)
        return n
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        n = once()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps
    return n * SZ / t / 1e9


for aligned in (True, False):
    for ns in (1, 8):
        print("aligned=%s streams=%d+%d both: %.1f GB/s per direction;  H2D alone: %.1f" % (
            aligned, ns, ns, run(aligned, ns), run(aligned, ns, both=False)), flush=True)
