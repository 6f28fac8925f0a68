"""Cross-GPU request fan-out (host plumbing over torch.distributed; NCCL over NVLink on the GPUs).

Connections shard across the GPUs of a box at accept time -- connection c terminates on GPU
``c mod N``, the mirror of the reference's round-robin over server pollsets
(src/core/lib/iomgr/tcp_server_posix.cc:255-258) -- and each connection's ring, cursors and credit
word are private to that GPU (pair.h:156-196), so the hot path itself needs no collective.  The one
exchange step appears when a server spreads *streams* over GPUs: a request deframed by ``k_recv`` on
the GPU that terminated its connection belongs to the GPU that owns the stream's handler.  Once per
poll epoch every rank hands over what it received for other owners in one grouped exchange:

    counts   all_to_all_single(int64[N])             how many requests / bytes go to each rank
    meta     all_to_all_single(int64[2 * requests])  (stream id, length) per request
    payload  all_to_all_single(uint8[bytes])          the request bytes, uneven splits

With the NCCL backend these are three ncclSend/ncclRecv groups over NVLink; other backends (gloo in
the CPU tests) lack all_to_all, so the same three steps run as batched isend / irecv.  Payload
tensors stay on the device they were delivered to; nothing is staged through the host.
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def owner_of_connection(conn_id: int, world: int) -> int:
    """GPU that terminates connection `conn_id` (accept-time round robin)."""
    return conn_id % world


def _all_to_all(out: torch.Tensor, inp: torch.Tensor, out_splits: Sequence[int], in_splits: Sequence[int], group):
    """all_to_all_single with uneven splits; batched send/recv where the backend has no all_to_all."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(out, inp, list(out_splits), list(in_splits), group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    outs = list(torch.split(out, list(out_splits)))
    ins = list(torch.split(inp, list(in_splits)))
    outs[rank].copy_(ins[rank])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if ins[peer].numel():
            ops.append(dist.P2POp(dist.isend, ins[peer].contiguous(), dist.get_global_rank(group, peer) if group else peer, group))
        if outs[peer].numel():
            ops.append(dist.P2POp(dist.irecv, outs[peer], dist.get_global_rank(group, peer) if group else peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class RequestFanout:
    """One grouped exchange per poll epoch.  `exchange` is collective: every rank calls it."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else torch.device("cpu")

    def exchange(self, requests: Sequence[Tuple[int, int, torch.Tensor]]) -> List[Tuple[int, int, torch.Tensor]]:
        """requests: (owner rank, stream id, uint8 payload) delivered on THIS rank this epoch.
        Returns the (source rank, stream id, payload) this rank owns, ordered by source rank and, within a
        source, in the order that source delivered them (per-stream order is preserved)."""
        W, dev = self.world, self.device
        by_dst = [[] for _ in range(W)]
        for owner, stream, payload in requests:
            if not 0 <= owner < W:
                raise ValueError("owner rank %d outside the group of %d" % (owner, W))
            by_dst[owner].append((stream, payload.reshape(-1)))
        send_cnt = torch.tensor([[len(v), sum(int(p.numel()) for _, p in v)] for v in by_dst], dtype=torch.int64,
                                device=dev).reshape(-1)
        recv_cnt = torch.empty_like(send_cnt)
        _all_to_all(recv_cnt, send_cnt, [2] * W, [2] * W, self.group)
        send_cnt_h, recv_cnt_h = send_cnt.reshape(W, 2).tolist(), recv_cnt.reshape(W, 2).tolist()
        # ---- metadata
        meta_out = torch.tensor([x for v in by_dst for (s, p) in v for x in (s, int(p.numel()))] or [0],
                                dtype=torch.int64, device=dev)[:2 * sum(c[0] for c in send_cnt_h)]
        meta_in = torch.empty(2 * sum(c[0] for c in recv_cnt_h), dtype=torch.int64, device=dev)
        _all_to_all(meta_in, meta_out, [2 * c[0] for c in recv_cnt_h], [2 * c[0] for c in send_cnt_h], self.group)
        # ---- payload
        parts = [p for v in by_dst for (_, p) in v]
        pay_out = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=dev)
        pay_in = torch.empty(sum(c[1] for c in recv_cnt_h), dtype=torch.uint8, device=dev)
        _all_to_all(pay_in, pay_out, [c[1] for c in recv_cnt_h], [c[1] for c in send_cnt_h], self.group)
        # ---- unpack
        out, mi, pi = [], 0, 0
        meta_h = meta_in.tolist()
        for src in range(W):
            for _ in range(recv_cnt_h[src][0]):
                stream, n = meta_h[mi], meta_h[mi + 1]
                mi += 2
                out.append((src, stream, pay_in[pi:pi + n]))
                pi += n
        return out


class InboxFanout:
    """The data-plane alternative of SURVEY section 8(e): instead of NCCL, a request deframed on GPU a for a stream
    owned by GPU b travels through an INBOX connection a -> b on the CUDA-IPC / NVLink wire -- the same ring-buffer
    frames, written by k_send straight into a ring in b's HBM and deframed there by k_recv (credit flows back over
    NVLink), i.e. through pairs of the C ABI (include/b200_pair.h): nothing but the bootstrap (48-byte address blobs,
    here over torch.distributed's object collectives) involves the host plumbing of RequestFanout.

    One inbox per ordered (source, destination) pair of ranks.  `exchange` is collective; per epoch each source sends,
    per destination, one block of (stream id, length) records followed by the payloads; the byte counts are
    exchanged first so that every rank knows what its inboxes will hold."""

    def __init__(self, pkg, device, group=None):
        self.pkg, self.group, self.dev = pkg, group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.tx = {r: pkg.Pair("inbox-%d-to-%d" % (self.rank, r)) for r in range(self.world) if r != self.rank}
        self.rx = {r: pkg.Pair("inbox-%d-from-%d" % (self.rank, r)) for r in range(self.world) if r != self.rank}
        mine = {"tx": {r: p.address() for r, p in self.tx.items()}, "rx": {r: p.address() for r, p in self.rx.items()}}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        for r in self.tx:
            if not self.tx[r].connect(everyone[r]["rx"][self.rank]) or not self.rx[r].connect(everyone[r]["tx"][self.rank]):
                raise RuntimeError("inbox connect failed: %s / %s" % (self.tx[r].error(), self.rx[r].error()))
        dist.barrier(group=group)

    def exchange(self, requests: Sequence[Tuple[int, int, torch.Tensor]]) -> List[Tuple[int, int, torch.Tensor]]:
        """Same contract as RequestFanout.exchange."""
        pkg, W, me = self.pkg, self.world, self.rank
        by_dst = [[] for _ in range(W)]
        for owner, stream, payload in requests:
            by_dst[owner].append((stream, payload.reshape(-1)))
        keep, sops, totals = [], [], {}
        for dst in range(W):
            if dst == me or not by_dst[dst]:
                continue
            meta = torch.tensor([x for s, p in by_dst[dst] for x in (s, int(p.numel()))], dtype=torch.int64,
                                device=self.dev).view(torch.uint8)
            sl = [(meta.data_ptr(), meta.numel())] + [(p.data_ptr(), p.numel()) for _, p in by_dst[dst] if p.numel()]
            arr = pkg.make_slices(sl)
            keep += [meta, arr]
            sops.append((self.tx[dst], arr, len(sl), 0))
            totals[dst] = sum(n for _, n in sl)
        counts = [None] * W
        dist.all_gather_object(counts, {d: (len(by_dst[d]), totals.get(d, 0)) for d in range(W)}, group=self.group)
        if sops:
            bs = pkg.Batch("send", sops, pkg.UNTIL_BLOCKED)
            bs.launch(None)
            got = bs.results(None)
            bs.destroy()
            if got != [totals[d] for d in sorted(totals)]:
                raise RuntimeError("inbox full: an epoch must fit half the inbox ring -- credit comes back in C/2 steps "
                                   "(%s of %s accepted)" % (got, totals))
        dist.barrier(group=self.group)  # the frames have landed in the destinations' rings
        rops, bufs = [], {}
        for src in range(W):
            if src == me:
                continue
            nreq, nbytes = counts[src].get(me, (0, 0))
            if nbytes:
                bufs[src] = (nreq, torch.empty(nbytes, dtype=torch.uint8, device=self.dev))
                rops.append((self.rx[src], bufs[src][1].data_ptr(), nbytes))
        if rops:
            br = pkg.Batch("recv", rops, pkg.UNTIL_BLOCKED)
            br.launch(None)
            got = br.results(None)
            br.destroy()
            if got != [bufs[s][1].numel() for s in sorted(bufs)]:
                raise RuntimeError("inbox delivered %s, expected %s" % (got, [bufs[s][1].numel() for s in sorted(bufs)]))
        res = []
        for src in range(W):
            if src == me:
                res += [(me, s, p) for s, p in by_dst[me]]
            elif src in bufs:
                nreq, buf = bufs[src]
                meta = buf[:16 * nreq].view(torch.int64).tolist()
                pos = 16 * nreq
                for k in range(nreq):
                    s, n = meta[2 * k], meta[2 * k + 1]
                    res.append((src, s, buf[pos:pos + n]))
                    pos += n
        return res

    def close(self):
        for p in list(self.tx.values()) + list(self.rx.values()):
            p.disconnect()
