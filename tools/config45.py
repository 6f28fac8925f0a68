"""BASELINE configs[3] and configs[4] on N GPUs of one box (run under torchrun, one rank per GPU).

  configs[3]  streaming payload sweep 1 KiB .. 16 MiB, 256 connections sharded over the GPUs, through the drop-in
              surface (b200_endpoint_write / read + b200_engine_work, batching engine, service kernels): every
              round of the > ring / > staging sizes (partial writes, credit returns) is driven from C
              (tools/native/ep_stream.cc), every delivered byte is compared; plus the NCCL request fan-out
              (grpc-rdma_b200/fanout.py) re-homing a quarter of the deframed requests to other GPUs.
  configs[4]  mixed 50/50: 512 connections per GPU (4096 on 8 GPUs), half unary 1 KiB ping-pong
              (tools/native/pingpong.cc), half streaming 4 MiB messages, AT THE SAME TIME through the same resident
              service kernels; beside it the reference's own CPU path (PairPollable ping-pong + the same streaming
              driver over the reference pair ops) for one rank's share of the connections on all host cores.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config45.py [--quick]

Rank 0 prints one JSON object (kept under profiles/)."""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench

quick = "--quick" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
numa = bench.bind_to_gpu_numa(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cpu_group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    # reductions / barriers WHILE THE SERVICE RUNS go over gloo on CPU tensors: a first-time CUDA kernel launch
    # (NCCL's or torch's) loads its module lazily, and that load can wait for an idle device -- which never comes
    # beside resident kernels
    cpu_group = dist.new_group(backend="gloo")
pkg = ge.load_package()
L = pkg.lib()
pkg.init(local)
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 16384)
libdir = os.path.dirname(pkg.LIB_PATH)
C.CDLL(pkg.ENDPOINT_LIB_PATH, mode=C.RTLD_GLOBAL)
ES = C.CDLL(os.path.join(libdir, "libb200_epstream.so"))
ES.ep_stream_run.restype = C.c_double
ES.ep_stream_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
PP = C.CDLL(os.path.join(libdir, "libb200_pingpong.so"))
PP.b200_pp_run.restype = C.c_double
PP.b200_pp_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]


T0 = time.time()


def log(*a):
    if rank == 0:
        print("[config45 %.1fs]" % (time.time() - T0), *a, file=sys.stderr, flush=True)


def allsum(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, group=cpu_group)
    return float(t.item())


def allmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=cpu_group)
    return float(t.item())


def barrier():
    # NOT dist.barrier(): its wait ends in a device-wide synchronisation, which never returns while the resident
    # service kernels run; an all_reduce + .item() only waits for its own stream
    if world > 1:
        allsum(0.0)


out = {"n_gpus": world, "numa": numa}
# request fan-out (NCCL all_to_all over NVLink): a quarter of the deframed requests belong to another GPU
if world > 1:
    import importlib
    fanout = importlib.import_module("grpc_rdma_b200.fanout")
    fx = fanout.RequestFanout(device=dev)
    rows = []
    for m in (1024, 65536, 4 << 20):
        n_req = max(8, min(1024, (1 << 27) // m))
        payload = torch.randint(0, 255, (n_req * m,), dtype=torch.uint8, device=dev)
        reqs, moved = [], 0
        for i in range(n_req):                                   # every 4th request belongs to another GPU
            owner = (rank + 1 + (i // 4) % (world - 1)) % world if i % 4 == 0 else rank
            moved += owner != rank
            reqs.append((owner, rank * n_req + i, payload[i * m:(i + 1) * m]))
        fx.exchange(reqs)                                        # warm-up
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            got = fx.exchange(reqs)
        e1.record()
        torch.cuda.synchronize()
        ms = allmax(e0.elapsed_time(e1) / 3)
        moved_bytes = allsum(float(moved * m))
        rows.append({"request_bytes": m, "requests_per_gpu": n_req, "rehomed_fraction": moved / n_req,
                     "exchange_ms": ms, "nvlink_GBps_all_gpus": moved_bytes / (ms * 1e-3) / 1e9,
                     "received_here": len(got)})
        del payload, reqs, got
    out["config4_fanout_nccl"] = rows
    log("fanout done")

assert L.b200_service_start(96) == 0, pkg.last_error()
log("service up")

# ------------------------------------------------------------------ configs[3]
conns4 = max(1, 256 // world)
sweep = []
sizes = [1024, 4096, 16384, 65536, 262144, 1 << 20, 4 << 20, 16 << 20]
for m in sizes:
    budget = (1 << 26) if quick else (1 << 28)                   # payload bytes per connection
    msgs = int(max(4, min(2000, budget // m)))
    o = (C.c_uint64 * 4)()
    barrier()
    t = ES.ep_stream_run(None, conns4, min(4, conns4), msgs, 2, m, 0, o)
    ok = t > 0 and o[1] == 0
    gbs = o[0] / t / 1e9 if t > 0 else 0.0
    row = {"message_bytes": m, "connections_per_gpu": conns4, "msgs_per_connection": msgs,
           "GBps_all_gpus": allsum(gbs), "msgs_per_s_all_gpus": allsum(conns4 * msgs / t if t > 0 else 0.0),
           "slowest_rank_s": allmax(t), "intact_all_ranks": allsum(0.0 if ok else 1.0) == 0.0}
    if m > (8 << 20):
        row["note"] = "message > ring (16 MiB - 24) and > staging (8 MiB): partial writes + C/2 credit returns, rounds driven from C"
    sweep.append(row)
    log("sweep", m, "->", round(row["GBps_all_gpus"], 2), "GB/s", round(row["msgs_per_s_all_gpus"]), "msgs/s", "t", round(t, 3))
out["config4_sweep_through_endpoint"] = sweep

# ------------------------------------------------------------------ configs[4]
conns5 = 128 if quick else 256                                   # unary and streaming each, per GPU
res = {}


def unary():
    it = 200 if quick else 400
    rtt = np.zeros(conns5 * it, dtype=np.uint64)
    t = PP.b200_pp_run(conns5, 8, it, 20, 1024, rtt.ctypes.data_as(C.POINTER(C.c_uint64)))
    r = np.sort(rtt) / 1e3
    res["unary"] = {"t": t, "p50": float(r[len(r) // 2]), "p99": float(r[int(len(r) * 0.99)]), "rt_per_s": conns5 * it / t if t > 0 else 0.0}


def streaming():
    o = (C.c_uint64 * 4)()
    t = ES.ep_stream_run(None, conns5, 8, 8 if quick else 16, 2, 4 << 20, 0, o)
    res["stream"] = {"t": t, "GBps": o[0] / t / 1e9 if t > 0 else 0.0, "bad": int(o[1])}


barrier()
th = [threading.Thread(target=unary), threading.Thread(target=streaming)]
for x in th:
    x.start()
for x in th:
    x.join()
log("config5 gpu part done", res)
out["config5_mixed"] = {
    "connections_per_gpu": 2 * conns5, "connections_total": 2 * conns5 * world,
    "unary_p50_us_worst_rank": allmax(res["unary"]["p50"]), "unary_p99_us_worst_rank": allmax(res["unary"]["p99"]),
    "unary_round_trips_per_s_all_gpus": allsum(res["unary"]["rt_per_s"]),
    "streaming_GBps_all_gpus": allsum(res["stream"]["GBps"]), "streaming_intact_all_ranks": allsum(float(res["stream"]["bad"])) == 0.0,
    "note": "both halves run at the same time through the same resident kernels (owner warps + 96 pool CTAs per GPU)"}
L.b200_service_stop()
barrier()

# the reference's CPU path for one rank's share, on all host cores (rank 0; the other ranks are idle now)
if rank == 0 and "--no-cpu" not in sys.argv:
    try:
        import orlib
        eng = orlib.Ref(debug=False) if orlib.ref_available(debug=False) else None
        R = C.CDLL(os.path.join(ROOT, "tests", "native", "libref_pair_ops_rel.so"))
        R.ref_pair_ops.restype = C.c_void_p
        R.ref_ops_config.argtypes = [C.c_uint32]
        R.ref_ops_config(16384)
        cores = bench.usable_cpus()
        cres = {}

        def c_unary():
            t, rtt = eng.bench_pingpong(conns5, max(1, min(16, cores // 8)), 200, 20, 1024, 16384 * 1024)
            r = np.sort(np.asarray(rtt).reshape(-1)) / 1e3
            cres["unary"] = {"p50_us": float(r[len(r) // 2]), "p99_us": float(r[int(len(r) * 0.99)]), "round_trips_per_s": conns5 * 200 / t}

        def c_stream():
            o = (C.c_uint64 * 4)()
            t = ES.ep_stream_run(R.ref_pair_ops(), conns5, max(1, cores // 4), 4, 1, 4 << 20, 0, o)
            cres["stream"] = {"GBps": o[0] / t / 1e9 if t > 0 else None, "bad": int(o[1]), "thread_pairs": max(1, cores // 4)}

        tt = [threading.Thread(target=c_unary), threading.Thread(target=c_stream)]
        for x in tt:
            x.start()
        for x in tt:
            x.join()
        cres["cores"] = cores
        cres["connections"] = 2 * conns5
        out["config5_cpu_reference_one_rank_share"] = cres
    except Exception as exc:
        out["config5_cpu_reference_one_rank_share"] = {"error": repr(exc)}
if rank == 0:
    print(json.dumps(out), flush=True)
barrier()
if world > 1:
    dist.destroy_process_group()
