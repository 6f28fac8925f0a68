"""Cross-GPU streaming over the CUDA-IPC / NVLink wire (run under torchrun, one rank per GPU):
rank r's senders are connected to rank (r+1) % N's receivers, so every k_send stores its frames into
the ring in the NEXT GPU's HBM over NVLink and every k_recv returns credit the other way.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/nvlink_bench.py [conns] [steps]

Prints one JSON line on rank 0: payload GB/s per GPU, wire bytes/s of k_send against the measured
770 GB/s per-direction peer copy of this pool (B200_PROFILING.md), k_recv against HBM, integrity."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
conns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
msg, ring_kb = 4 * 1024 * 1024, 16384
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ["B200_IPC_WIRE"] = "1"
pkg = ge.load_package()
pkg.init(local)
L = pkg.lib()
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", ring_kb)
dev = torch.device("cuda", local)
tx = [pkg.Pair("nv-tx-%d-%d" % (rank, c)) for c in range(conns)]
rx = [pkg.Pair("nv-rx-%d-%d" % (rank, c)) for c in range(conns)]
mine = {"tx": [p.address() for p in tx], "rx": [p.address() for p in rx]}
everyone = [None] * world
dist.all_gather_object(everyone, mine)                      # the TCP bootstrap's job
nxt, prv = (rank + 1) % world, (rank - 1) % world
for c in range(conns):
    assert tx[c].connect(everyone[nxt]["rx"][c]), tx[c].error()
    assert rx[c].connect(everyone[prv]["tx"][c]), rx[c].error()
dist.barrier()
lens = pkg.chttp2_slice_lens(msg)
total = sum(lens)
tx_alg, rx_alg = pkg.frame_hbm_bytes(lens)
wire_bytes = sum(16 + (n + 7) // 8 * 8 for n in lens)

i = torch.arange(total, device=dev, dtype=torch.int64)
row = (((i * 2654435761) >> 11) & 255).to(torch.uint8)


def payload(r):
    offs = ((torch.arange(conns, device=dev, dtype=torch.int64) + r * conns) * 131 & 255).to(torch.uint8)
    return (row[None, :] + offs[:, None]).reshape(-1)


src, dst = payload(rank), torch.zeros(conns * total, dtype=torch.uint8, device=dev)
expect = payload(prv)
sops, rops, keep = [], [], []
for c in range(conns):
    off, sl = 0, []
    for n in lens:
        sl.append((src.data_ptr() + c * total + off, n))
        off += n
    arr = pkg.make_slices(sl)
    keep.append(arr)
    sops.append((tx[c], arr, len(lens), 0))
    rops.append((rx[c], dst.data_ptr() + c * total, total))
bs = pkg.Batch("send", sops, pkg.UNTIL_BLOCKED)
br = pkg.Batch("recv", rops, pkg.UNTIL_BLOCKED)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sh = C.c_void_p(stream.cuda_stream)
send_ms, recv_ms = [], []
for k in range(steps + 3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    dist.barrier()
    e[0].record(stream)
    bs.launch(sh)
    e[1].record(stream)
    stream.synchronize()
    dist.barrier()                                          # the neighbour's frames have landed in my rings
    e[2].record(stream)
    br.launch(sh)
    e[3].record(stream)
    stream.synchronize()
    if k >= 3:
        send_ms.append(e[0].elapsed_time(e[1]))
        recv_ms.append(e[2].elapsed_time(e[3]))
ok = bs.results(sh) == [total] * conns and br.results(sh) == [total] * conns and bool(torch.equal(dst, expect))
t = torch.tensor([sum(send_ms) / len(send_ms), sum(recv_ms) / len(recv_ms), 0.0 if ok else 1.0], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    s_ms, r_ms, bad = t.tolist()
    print(json.dumps({
        "what": "k_send over the CUDA-IPC/NVLink wire: frames stored into the next GPU's rings",
        "n_gpus": world, "connections_per_gpu": conns, "message_bytes": msg, "steps": steps, "intact": bad == 0.0,
        "k_send_ms": s_ms, "k_recv_ms": r_ms,
        "payload_GBps_per_gpu_send": conns * msg / (s_ms * 1e-3) / 1e9,
        "nvlink_write_GBps_per_gpu": conns * wire_bytes / (s_ms * 1e-3) / 1e9,
        "nvlink_peak_GBps": 770.0, "nvlink_frac": conns * wire_bytes / (s_ms * 1e-3) / 1e9 / 770.0,
        "k_recv_hbm_GBps": conns * rx_alg / (r_ms * 1e-3) / 1e9,
        "note": "time = max over ranks; every rank sends and receives at once, so each NVLink direction carries one stream"}),
        flush=True)
for p in tx + rx:
    p.disconnect()
dist.barrier()
dist.destroy_process_group()
