"""2+ GPUs: grpc_rdma_b200.fanout.InboxFanout -- the request fan-out through inbox connections on the CUDA-IPC /
NVLink wire (k_send into the owner GPU's ring, k_recv there), same check as tools/fanout_nccl_check.py.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/fanout_inbox_check.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ["B200_IPC_WIRE"] = "1"
pkg = ge.load_package()
pkg.init(local)
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 16384)
from importlib import import_module
fanout = import_module("grpc_rdma_b200.fanout")
dev = torch.device("cuda", local)
fo = fanout.InboxFanout(pkg, dev)


def payload(src, stream, n):
    i = torch.arange(n, device=dev, dtype=torch.int64)
    return ((i * 31 + src * 97 + stream * 13) & 255).to(torch.uint8)


ok = True
for epoch in range(3):
    reqs = [(s % world, s, payload(rank, s, (4099 * (s + 1) + epoch) % 70000)) for s in range(rank, 64, world)]
    got = fo.exchange(reqs)
    want = [(src, s) for src in range(world) for s in range(src, 64, world) if s % world == rank]
    ok = ok and [(a, b) for a, b, _ in got] == want
    for src, s, p in got:
        ok = ok and bool(torch.equal(p, payload(src, s, (4099 * (s + 1) + epoch) % 70000)))
# throughput: 1 MiB requests, every request owned by the next GPU (an epoch must fit half the inbox ring: credit
# comes back in C/2 steps)
big = [((rank + 1) % world, 1000 + k, payload(rank, k, 1 << 20)) for k in range(4)]
fo.exchange(big)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    got = fo.exchange(big)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
for src, s, p in got:
    ok = ok and bool(torch.equal(p, payload(src, s - 1000, 1 << 20)))
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("fanout through inbox connections on the NVLink wire, %d GPUs: %s; 4 x 1 MiB per GPU per epoch in %.2f ms (%.1f GB/s per GPU incl. the host-side epoch protocol)"
          % (world, "OK" if t.item() == 0 else "MISMATCH", dt * 1e3, 4 * (1 << 20) / dt / 1e9), flush=True)
fo.close()
dist.barrier()
dist.destroy_process_group()
