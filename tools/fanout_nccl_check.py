"""2+ GPUs, NCCL: grpc_rdma_b200.fanout.RequestFanout on device tensors (the same check as the gloo CPU test).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/fanout_nccl_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ge.load_package()
from importlib import import_module
fanout = import_module("grpc_rdma_b200.fanout")
dev = torch.device("cuda", local)
fo = fanout.RequestFanout(device=dev)


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
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("fanout over NCCL on %d GPUs: %s" % (world, "OK" if t.item() == 0 else "MISMATCH"), flush=True)
dist.destroy_process_group()
