"""BASELINE configs[3] on one GPU: streaming payload sweep 1 KiB .. 16 MiB, 256 connections, device-resident,
chttp2-shaped slices, 16 MiB rings.  A message larger than what the ring admits (16 MiB > C - 24, and more than
the C/2 a single Send accepts) goes through the partial-write / credit path: Send and Recv batches alternate
until everything is delivered, exactly like rdma_flush / rdma_do_read re-entered from the poll loop.
Prints one JSON line per size (experiment / documentation helper; profiles/)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
pkg.init(0)
L = pkg.lib()
conns, ring_kb = 256, 16384
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", ring_kb)
pkg.config_set("GRPC_RDMA_MAX_SGE", 30)
dev = torch.device("cuda", 0)
pairs = [pkg.connected_pair("sw-tx%d" % c, "sw-rx%d" % c) for c in range(conns)]
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sh = C.c_void_p(stream.cuda_stream)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0

for msg in [1 << k for k in (10, 12, 14, 16, 18, 20, 22, 24)]:
    lens = pkg.chttp2_slice_lens(msg)
    total = sum(lens)
    tx_alg, rx_alg = pkg.frame_hbm_bytes(lens)
    i = torch.arange(total, device=dev, dtype=torch.int64)
    row = (((i * 2654435761) >> 11) & 255).to(torch.uint8)
    offs = (torch.arange(conns, device=dev, dtype=torch.int64) * 131 & 255).to(torch.uint8)
    src = (row[None, :] + offs[:, None]).reshape(-1)
    dst = torch.zeros(conns * total, dtype=torch.uint8, device=dev)
    keep = []

    def one_message():
        """returns number of (send, recv) rounds"""
        done_s, done_r, rounds = [0] * conns, [0] * conns, 0
        # position bookkeeping on the host between rounds, like rdma_flush's cursor
        starts = [0]
        for n in lens:
            starts.append(starts[-1] + n)
        while min(done_r) < total:
            rounds += 1
            sops, rops = [], []
            for c in range(conns):
                if done_s[c] < total:
                    k = 0
                    while starts[k + 1] <= done_s[c]:
                        k += 1
                    arr = pkg.make_slices([(src.data_ptr() + c * total + starts[j], lens[j]) for j in range(k, len(lens))])
                    keep.append(arr)
                    sops.append((c, (pairs[c][0], arr, len(lens) - k, done_s[c] - starts[k])))
            if sops:
                bs = pkg.Batch("send", [o for _, o in sops], pkg.UNTIL_BLOCKED)
                bs.launch(sh)
                for (c, _), n in zip(sops, bs.results(sh)):
                    done_s[c] += n
                bs.destroy()
            for c in range(conns):
                if done_r[c] < total:
                    rops.append((c, (pairs[c][1], dst.data_ptr() + c * total + done_r[c], total - done_r[c])))
            br = pkg.Batch("recv", [o for _, o in rops], pkg.UNTIL_BLOCKED)
            br.launch(sh)
            for (c, _), n in zip(rops, br.results(sh)):
                done_r[c] += n
            br.destroy()
        return rounds

    if msg <= (4 << 20):
        # fits the ring: prepared batches, timed back to back like bench.py
        sops, rops = [], []
        for c in range(conns):
            off, sl = 0, []
            for n in lens:
                sl.append((src.data_ptr() + c * total + off, n))
                off += n
            arr = pkg.make_slices(sl)
            keep.append(arr)
            sops.append((pairs[c][0], arr, len(lens), 0))
            rops.append((pairs[c][1], dst.data_ptr() + c * total, total))
        bs, br = pkg.Batch("send", sops, pkg.UNTIL_BLOCKED), pkg.Batch("recv", rops, pkg.UNTIL_BLOCKED)
        reps = 10
        for _ in range(3):
            bs.launch(sh)
            br.launch(sh)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            bs.launch(sh)
            br.launch(sh)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ok = bs.results(sh) == [total] * conns and br.results(sh) == [total] * conns and bool(torch.equal(src, dst))
        rounds = 1
        bs.destroy()
        br.destroy()
    else:
        one_message()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rounds = one_message()
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)      # includes the host round trips between rounds
        ok = bool(torch.equal(src, dst))
    print(json.dumps({"message_bytes": msg, "slices": len(lens), "connections": conns, "ms_per_message_round": ms,
                      "payload_GBps": conns * msg / ms / 1e6, "msgs_per_s": conns / ms * 1e3,
                      "hbm_frac_of_measured_peak": conns * (tx_alg + rx_alg) / ms / 1e6 / peak, "rounds": rounds,
                      "intact": ok}), flush=True)
    del src, dst
    keep.clear()
