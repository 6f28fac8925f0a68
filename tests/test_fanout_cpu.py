"""CPU, world_size 2, gloo: the host logic of the multi-GPU path -- connection sharding and the
grouped cross-rank request fan-out (grpc-rdma_b200/fanout.py).  On the GPUs the same three exchange
steps are NCCL all_to_all over NVLink."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _payload(src, stream, k, n):
    i = np.arange(n, dtype=np.uint64)
    return torch.from_numpy(((i * np.uint64(31) + np.uint64(src * 97 + stream * 13 + k)) & np.uint64(255)).astype(np.uint8))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from importlib import import_module
    fanout = import_module("grpc_rdma_b200.fanout")
    fo = fanout.RequestFanout()
    ok = True
    # connections shard by accept order; streams are owned by stream_id mod world
    conns = [c for c in range(16) if fanout.owner_of_connection(c, world) == rank]
    ok = ok and conns == list(range(rank, 16, world))
    for epoch in range(3):
        reqs, expect_local = [], []
        for c in conns:
            for k in range(2 + (c + epoch) % 3):
                stream = c * 8 + k
                n = (37 * (c + 1) * (k + 1) + epoch) % 5000          # includes empty and odd sizes
                reqs.append((stream % world, stream, _payload(rank, stream, epoch, n)))
        got = fo.exchange(reqs)
        # what this rank must hold now: every request of every rank whose stream it owns, grouped by source
        want = []
        for src in range(world):
            for c in range(src, 16, world):
                for k in range(2 + (c + epoch) % 3):
                    stream = c * 8 + k
                    if stream % world == rank:
                        n = (37 * (c + 1) * (k + 1) + epoch) % 5000
                        want.append((src, stream, _payload(src, stream, epoch, n)))
        ok = ok and len(got) == len(want)
        for (gs, gst, gp), (ws, wst, wp) in zip(got, want):
            ok = ok and gs == ws and gst == wst and torch.equal(gp, wp)
    # an epoch in which one rank has nothing to hand over
    got = fo.exchange([] if rank == 0 else [(0, 5, _payload(rank, 5, 9, 100))])
    ok = ok and (len(got) == (world - 1) if rank == 0 else len(got) == 0)
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_fanout_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
