"""One end of a cross-GPU connection (tests/test_ipc_wire_gpu.py and tools): one process per GPU, the
48-byte address blobs are exchanged through files (the TCP bootstrap's job in gRPC), everything
else goes GPU -> NVLink -> GPU through the CUDA-IPC wire of libb200rdma.so.

    python ipc_wire_worker.py <role: client|server> <device> <dir> <ring_kb> <msg_bytes> <n_msgs> [conns]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge


def pattern(c, m, n):
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761) >> np.uint64(11)) + np.uint64(131 * c + 7 * m)).astype(np.uint8)


def wait_file(path, timeout=120):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.01)
    time.sleep(0.02)
    return open(path, "rb").read()


def put_file(path, data):
    with open(path + ".tmp", "wb") as f:
        f.write(data)
    os.rename(path + ".tmp", path)


def main():
    role, dev, d, ring_kb, msg, n_msgs = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    conns = int(sys.argv[7]) if len(sys.argv) > 7 else 1
    os.environ["B200_IPC_WIRE"] = "1"
    pkg = ge.load_package()
    pkg.init(dev)
    L = pkg.lib()
    pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", ring_kb)
    me, other = ("c", "s") if role == "client" else ("s", "c")
    pairs = [pkg.Pair("%s%d" % (me, c)) for c in range(conns)]
    for c, p in enumerate(pairs):
        put_file(os.path.join(d, "%s%d.addr" % (me, c)), p.address())
    for c, p in enumerate(pairs):
        assert p.connect(wait_file(os.path.join(d, "%s%d.addr" % (other, c)))), p.error()
    if os.environ.get("IPC_WIRE_MODE") == "death":
        # liveness: the server dies without Disconnect right after connecting; the client must see HALF_CLOSED
        if role == "server":
            wait_file(os.path.join(d, "client.connected"))
            os._exit(0)
        put_file(os.path.join(d, "client.connected"), b"1")
        t0 = time.time()
        while pairs[0].status() != 3 and time.time() - t0 < 20:
            time.sleep(0.05)
        put_file(os.path.join(d, "client.json"), json.dumps({"half_closed": pairs[0].status() == 3,
                                                            "seconds": time.time() - t0}).encode())
        os._exit(0)
    lens = pkg.chttp2_slice_lens(msg)
    total = sum(lens)
    res = {"role": role, "ok": True, "conns": conns}
    buf = L.b200_mem_alloc_device(conns * total)
    host = np.zeros(conns * total, dtype=np.uint8)

    def slices_for(c):
        off, sl = 0, []
        for n in lens:
            sl.append((buf + c * total + off, n))
            off += n
        return pkg.make_slices(sl)

    if role == "client":
        t_send = 0.0
        for m in range(n_msgs):
            for c in range(conns):
                host[c * total:(c + 1) * total] = pattern(c, m, total)
            assert L.b200_memcpy(buf, host.ctypes.data, host.size, 0, None) == 0 and L.b200_stream_sync(None) == 0
            keep = [slices_for(c) for c in range(conns)]
            done = [0] * conns          # bytes accepted so far per connection
            t0 = time.perf_counter()
            while min(done) < total:
                ops, idxs = [], []
                for c in range(conns):
                    if done[c] >= total:
                        continue
                    # position (slice index, byte index) of done[c] in the slice list
                    acc, i = 0, 0
                    while acc + lens[i] <= done[c]:
                        acc += lens[i]
                        i += 1
                    sub = pkg.make_slices([(keep[c][j].ptr, keep[c][j].len) for j in range(i, len(lens))])
                    keep.append(sub)
                    ops.append((pairs[c], sub, len(lens) - i, done[c] - acc))
                    idxs.append(c)
                bt = pkg.Batch("send", ops, pkg.UNTIL_BLOCKED)
                bt.launch(None)
                r = bt.results(None)
                bt.destroy()
                for c, n in zip(idxs, r):
                    done[c] += n
                if not any(r):
                    time.sleep(0.0005)      # ring full: wait for credit from the other GPU
            t_send += time.perf_counter() - t0
        res["send_seconds"] = t_send
        res["bytes"] = n_msgs * conns * total
        # half close: the server must see peer_exit
        wait_file(os.path.join(d, "server.done"))
        for p in pairs:
            p.disconnect()
    else:
        ok = True
        for m in range(n_msgs):
            got = [0] * conns
            t0 = time.time()
            while min(got) < total:
                ops = [(pairs[c], buf + c * total + got[c], total - got[c]) for c in range(conns) if got[c] < total]
                idxs = [c for c in range(conns) if got[c] < total]
                bt = pkg.Batch("recv", ops, pkg.UNTIL_BLOCKED)
                bt.launch(None)
                r = bt.results(None)
                bt.destroy()
                for c, n in zip(idxs, r):
                    got[c] += n
                if time.time() - t0 > 120:
                    raise TimeoutError("message %d: got %s of %d" % (m, got, total))
            assert L.b200_memcpy(host.ctypes.data, buf, host.size, 1, None) == 0 and L.b200_stream_sync(None) == 0
            for c in range(conns):
                ok = ok and bool(np.array_equal(host[c * total:(c + 1) * total], pattern(c, m, total)))
        res["ok"] = ok
        st = pairs[0].state()
        res["ring_empty"] = bool(not pairs[0].ring_image().any())
        res["state"] = st
        put_file(os.path.join(d, "server.done"), b"1")
        t0 = time.time()
        while pairs[0].status() != 3 and time.time() - t0 < 30:   # HALF_CLOSED once the client left
            time.sleep(0.01)
        res["half_closed"] = pairs[0].status() == 3
        for p in pairs:
            p.disconnect()
    put_file(os.path.join(d, role + ".json"), json.dumps(res).encode())


if __name__ == "__main__":
    main()
