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
