"""PCIe calibration 3: which alignment matters for bidirectional copies? (experiment helper)"""
import time

import torch

N, SZ = 128, 4196622
PAD = 8192
h_src = torch.empty(N * (SZ + PAD) + 65536, dtype=torch.uint8).pin_memory()
h_dst = torch.empty(N * (SZ + PAD) + 65536, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N * (SZ + PAD) + 65536, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N * (SZ + PAD) + 65536, dtype=torch.uint8, device="cuda")
ups = [torch.cuda.Stream() for _ in range(8)]
downs = [torch.cuda.Stream() for _ in range(8)]
base_h = (-h_src.data_ptr()) % 4096
base_h2 = (-h_dst.data_ptr()) % 4096
base_d = (-d_a.data_ptr()) % 4096
base_d2 = (-d_b.data_ptr()) % 4096


def run(hoff, doff, size, tiny=0, reps=3):
    """copy c: host offset = c*stride + hoff, device offset = c*stride + doff (stride 4 KiB multiple)"""
    stride = ((SZ + PAD) // 4096) * 4096

    def once():
        for c in range(N):
            ho, do = c * stride + hoff, c * stride + doff
            with torch.cuda.stream(ups[c % 8]):
                d_a[base_d + do:base_d + do + size].copy_(h_src[base_h + ho:base_h + ho + size], non_blocking=True)
                for t in range(tiny):
                    d_a[base_d + do + size + 64 * t:base_d + do + size + 64 * t + 33].copy_(
                        h_src[base_h + ho + size + 64 * t:base_h + ho + size + 64 * t + 33], non_blocking=True)
            with torch.cuda.stream(downs[c % 8]):
                h_dst[base_h2 + ho:base_h2 + ho + size].copy_(d_b[base_d2 + do:base_d2 + do + size], non_blocking=True)
                for t in range(tiny):
                    h_dst[base_h2 + ho + size + 64 * t:base_h2 + ho + size + 64 * t + 33].copy_(
                        d_b[base_d2 + do + size + 64 * t:base_d2 + do + size + 64 * t + 33], non_blocking=True)
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return N * size / ((time.perf_counter() - t0) / reps) / 1e9


for name, hoff, doff, size, tiny in [
        ("all aligned, odd size", 0, 0, SZ, 0),
        ("all aligned, 4K-multiple size", 0, 0, SZ // 4096 * 4096, 0),
        ("host+2 dev+2 (co-aligned)", 2, 2, SZ, 0),
        ("host+64 dev+64", 64, 64, SZ, 0),
        ("host+256 dev+256", 256, 256, SZ, 0),
        ("host+2 dev aligned", 2, 0, SZ, 0),
        ("host aligned dev+2", 0, 2, SZ, 0),
        ("host+1366 dev+1366", 1366, 1366, SZ, 0),
        ("all aligned + 2 tiny copies each", 0, 0, SZ // 4096 * 4096, 2)]:
    print("%-36s %.1f GB/s per direction" % (name, run(hoff, doff, size, tiny)), flush=True)
