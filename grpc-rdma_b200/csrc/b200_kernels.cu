// b200_kernels.cu -- sm_100a kernels of the RDMA_BPEV endpoint hot path.
//
//   k_send       gather/encode: grpc_slice list -> [len][payload][pad][~0] frames
//                written straight at the remote tail of the peer's HBM ring
//                (replaces PairPollable::Send pair.cc:645-734 + AppendHeader/
//                Payload/Footer ring_buffer.h:84-99 + GetWriteRequests
//                ring_buffer.cc:261-330 + the NIC's RDMA write)
//   k_recv       deframe/scatter + clear-on-read + credit write-back (replaces
//                RingBufferPollable::Read ring_buffer.cc:122-191 and
//                PairPollable::Recv/updateStatus pair.cc:264-286,624-641)
//   k_poll_scan  readiness scan (replaces the per-pair body of
//                Poller::begin_polling poller.cc:66-101 and of the engine's
//                busy-poll window ev_epollex_rdma_bpev_linux.cc:1104-1145)
//
// Pure indexing / memcpy work: HBM-bound, no tensor cores.  All bulk traffic is
// 16-byte vector loads/stores; byte granularity only at the <16-byte edges of
// a copy.  One CTA serves one (pair, op); inside the CTA, warp 0 does the
// per-call integer planning with warp scans and all warps move bytes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200_dev.cuh"

namespace b200 {

// ------------------------------------------------------------ memory helpers

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// ring reads must not use the non-coherent path: the ring is written by other
// kernels / the wire while we run
__device__ __forceinline__ uint4 ld_ring16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_u64(const void* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_v2u64(void* p, uint64_t a, uint64_t b) {
  // 16-byte status_report {remote_head, peer_exit}: fence + one vector store
  __threadfence_system();
  asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

// Take bytes [m, m+16) out of the 32-byte concatenation A|B (m = 1..15).
template <bool kHi, bool kShift>
__device__ __forceinline__ uint4 shift_window(uint4 A, uint4 B, unsigned sh) {
  uint64_t a0 = (uint64_t)A.x | ((uint64_t)A.y << 32), a1 = (uint64_t)A.z | ((uint64_t)A.w << 32);
  uint64_t b0 = (uint64_t)B.x | ((uint64_t)B.y << 32), b1 = (uint64_t)B.z | ((uint64_t)B.w << 32);
  uint64_t w0 = kHi ? a1 : a0, w1 = kHi ? b0 : a1, w2 = kHi ? b1 : b0;
  uint64_t lo, hi;
  if (kShift) {
    lo = (w0 >> sh) | (w1 << (64 - sh));
    hi = (w1 >> sh) | (w2 << (64 - sh));
  } else {
    lo = w0;
    hi = w1;
  }
  uint4 r;
  r.x = (uint32_t)lo;
  r.y = (uint32_t)(lo >> 32);
  r.z = (uint32_t)hi;
  r.w = (uint32_t)(hi >> 32);
  return r;
}

template <bool kRingSrc>
__device__ __forceinline__ uint4 ld_src16(const void* p) {
  return kRingSrc ? ld_ring16(p) : ld_stream16(p);
}

constexpr int kUnroll = 8;  // 16-byte vectors in flight per lane: one 4 KiB work item = one block

__device__ __forceinline__ uint4 shfl_down1_wrap(uint4 a, uint4 next0, uint32_t lane) {
  // lane L gets lane L+1's vector; lane 31 gets lane 0's vector of the next row
  uint4 t, w;
  t.x = __shfl_down_sync(0xffffffffu, a.x, 1);
  t.y = __shfl_down_sync(0xffffffffu, a.y, 1);
  t.z = __shfl_down_sync(0xffffffffu, a.z, 1);
  t.w = __shfl_down_sync(0xffffffffu, a.w, 1);
  w.x = __shfl_sync(0xffffffffu, next0.x, 0);
  w.y = __shfl_sync(0xffffffffu, next0.y, 0);
  w.z = __shfl_sync(0xffffffffu, next0.z, 0);
  w.w = __shfl_sync(0xffffffffu, next0.w, 0);
  return lane == 31 ? w : t;
}

// Warp-cooperative: dst 16-byte aligned, src misaligned by m = src & 15 (1..15).  Output vector i
// is cut out of aligned source vectors i and i+1; vector i+1 is the neighbour lane's load (shuffle),
// so every source byte is fetched once and kUnroll loads per lane are in flight.  Source vector
// `nvec` always contains a byte of the range (never faults); vectors beyond it are not touched.
// One block = up to 32*kUnroll output vectors (4 KiB).  All kUnroll loads of a lane are in
// flight at once; a short block is handled by predication, never by a slower loop.  Source
// word w exists iff w <= nvec, output vector v exists iff v < nvec (nvec <= 32*kUnroll).
template <bool kRingSrc, bool kHi, bool kShift>
__device__ __noinline__ void copy_block_shifted(uint4* __restrict__ d, const uint4* __restrict__ s, uint32_t nvec,
                                                unsigned sh, uint32_t lane) {
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const uint4* sl = s + lane;
  uint4* dl = d + lane;
  uint4 A[kUnroll + 1];
#pragma unroll
  for (int k = 0; k <= kUnroll; k++) A[k] = (32u * k + lane <= nvec) ? ld_src16<kRingSrc>(sl + 32 * k) : zero;
#pragma unroll
  for (int k = 0; k < kUnroll; k++) {
    const uint4 B = shfl_down1_wrap(A[k], A[k + 1], lane);
    if (32u * k + lane < nvec) st_stream16(dl + 32 * k, shift_window<kHi, kShift>(A[k], B, sh));
  }
}

template <bool kRingSrc>
__device__ __noinline__ void copy_block_aligned(uint4* __restrict__ d, const uint4* __restrict__ s, uint32_t nvec,
                                                uint32_t lane) {
  const uint4* sl = s + lane;
  uint4* dl = d + lane;
  uint4 v[kUnroll];
#pragma unroll
  for (int k = 0; k < kUnroll; k++)
    if (32u * k + lane < nvec) v[k] = ld_src16<kRingSrc>(sl + 32 * k);
#pragma unroll
  for (int k = 0; k < kUnroll; k++)
    if (32u * k + lane < nvec) st_stream16(dl + 32 * k, v[k]);
}

// Warp-cooperative copy of n bytes, any alignment on either side.  Bulk = aligned 16-byte
// stores; the <16-byte head and tail use one byte per lane.
template <bool kRingSrc>
__device__ __forceinline__ uint8_t ld_byte(const uint8_t* p) {
  return kRingSrc ? *reinterpret_cast<const volatile uint8_t*>(p) : __ldg(p);
}

template <bool kRingSrc>
__device__ __forceinline__ void coop_copy(uint8_t* dst, const uint8_t* src, uint64_t n, uint32_t lane) {
  if (n == 0) return;
  uint64_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head > n) head = n;
  const uint64_t nvec = (n - head) >> 4;
  const uint64_t tail = n - head - (nvec << 4);
  // <16-byte edges: one byte per lane
  if (lane < head) dst[lane] = ld_byte<kRingSrc>(src + lane);
  if (lane < tail) dst[head + (nvec << 4) + lane] = ld_byte<kRingSrc>(src + head + (nvec << 4) + lane);
  const uint8_t* vsrc = src + head;
  const unsigned m = (unsigned)(reinterpret_cast<uintptr_t>(vsrc) & 15);
  constexpr uint64_t kBlk = 32ull * kUnroll;
  const unsigned sh = (m & 7) * 8;
  const uint4* sal = reinterpret_cast<const uint4*>(vsrc - m);
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  for (uint64_t v0 = 0; v0 < nvec; v0 += kBlk) {
    const uint32_t nb = (uint32_t)(nvec - v0 < kBlk ? nvec - v0 : kBlk);
    if (m == 0) copy_block_aligned<kRingSrc>(dv + v0, sal + v0, nb, lane);
    else if (!(m & 8)) copy_block_shifted<kRingSrc, false, true>(dv + v0, sal + v0, nb, sh, lane);
    else if (sh) copy_block_shifted<kRingSrc, true, true>(dv + v0, sal + v0, nb, sh, lane);
    else copy_block_shifted<kRingSrc, true, false>(dv + v0, sal + v0, nb, sh, lane);
  }
}

// Warp-cooperative zero fill of n bytes at p (any alignment).
__device__ __forceinline__ void coop_zero(uint8_t* p, uint64_t n, uint32_t lane) {
  if (n == 0) return;
  uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
  if (head > n) head = n;
  if (lane < head) p[lane] = 0;
  p += head;
  n -= head;
  const uint64_t nvec = n >> 4;
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint64_t i = lane; i < nvec; i += 32) st_stream16(d + i, z);
  const uint64_t done = nvec << 4;
  if (lane < n - done) p[done + lane] = 0;
}

// GetReadableSize / HasMessage of a pair whose cursor is (head, remain)
// (ring_buffer.cc:56-97).  A header larger than cap-24 is a torn read in the
// reference (it spins); here it reports "not readable yet".
__device__ __forceinline__ void rx_probe(const uint8_t* ring, uint64_t cap, uint64_t head, uint64_t remain,
                                         uint32_t& has_msg, uint64_t& readable) {
  if (remain > 0) {
    has_msg = 1;
    readable = remain;
    return;
  }
  uint64_t hdr = ld_acquire_u64(ring + head);
  has_msg = hdr != 0;
  readable = 0;
  if (hdr != 0 && hdr <= cap - kReserved) {
    uint64_t foot = ld_acquire_u64(ring + ((head + 8 + round_up8(hdr)) & (cap - 1)));
    if (foot == kFooter) readable = hdr;
  }
}

__device__ __forceinline__ void publish_mirror(PairMirror* m, const PairDev* P, uint32_t has_msg,
                                               uint64_t readable) {
  if (m == nullptr) return;
  volatile PairMirror* vm = m;
  vm->head = P->head;
  vm->moving_head = P->moving_head;
  vm->remain = P->remain;
  vm->acc = P->acc;
  vm->remote_tail = P->remote_tail;
  vm->credit_head = P->credit_head;
  vm->readable = readable;
  vm->partial_write = P->partial_write;
  vm->peer_exit = P->credit_exit;
  vm->has_message = has_msg;
  __threadfence_system();
  vm->seq = vm->seq + 1;
}

// =========================================================================
// Producer / consumer skeleton shared by k_send and k_recv
// =========================================================================
//
// One CTA per (pair, op).  Warp 0 is the producer: it runs the reference's
// integer logic (Send planning / frame-list walking) ahead of the data movers
// and publishes 4 KiB work items into a ticket ring in shared memory.  All
// other warps (and warp 0 once it has nothing left to publish) claim items from
// a shared counter and move bytes.  There is no CTA barrier on the steady-state
// path; a "segment" ends only where the protocol needs everything before it to
// be finished: the footer flush of Send, the credit write of Recv, the end of
// the op.

constexpr int kThreads = 512;
constexpr uint32_t kChunk = 4096;  // payload bytes per work item
constexpr uint32_t kQI = 256;      // ticket ring entries (1 MiB of look-ahead)

struct WorkItem {      // 32 bytes
  uint64_t a;          // send: source pointer          recv: ring offset of the payload bytes
  uint64_t b;          // send: ring offset (payload)   recv: offset in the destination
  uint64_t c;          // send: header value (chunk 0)  recv: zhead | ztail << 16
  uint32_t n;          // bytes
  uint32_t ready;      // ticket: item id + 1 when published, 0 when free
};

struct PipeCtl {
  uint32_t next;         // next item id to claim
  uint32_t total_items;  // valid once seg_done
  uint32_t seg_done;
  uint32_t op_done;
};

__device__ __forceinline__ uint32_t ld_shared_volatile(const uint32_t* p) { return *(const volatile uint32_t*)p; }

// producer side: wait for the slot of item `id`, fill it, publish
__device__ __forceinline__ void publish_item(WorkItem* q, uint32_t id, uint64_t a, uint64_t b, uint64_t c,
                                             uint32_t n) {
  WorkItem* slot = &q[id % kQI];
  while (ld_shared_volatile(&slot->ready) != 0) __nanosleep(20);
  slot->a = a;
  slot->b = b;
  slot->c = c;
  slot->n = n;
  __threadfence_block();
  *(volatile uint32_t*)&slot->ready = id + 1;
}

// consumer side: claim the next item; false when the segment is drained
__device__ __forceinline__ bool claim_item(WorkItem* q, PipeCtl* ctl, uint32_t lane, uint64_t& a, uint64_t& b,
                                           uint64_t& c, uint32_t& n) {
  uint32_t w = 0;
  if (lane == 0) w = atomicAdd(&ctl->next, 1u);
  w = __shfl_sync(0xffffffffu, w, 0);
  WorkItem* slot = &q[w % kQI];
  bool got = false;
  while (true) {
    if (ld_shared_volatile(&slot->ready) == w + 1) {
      got = true;
      break;
    }
    if (ld_shared_volatile(&ctl->seg_done) && w >= ld_shared_volatile(&ctl->total_items)) {
      // re-check: the item may have been published between the two reads
      got = ld_shared_volatile(&slot->ready) == w + 1;
      break;
    }
    __nanosleep(40);
  }
  got = __any_sync(0xffffffffu, got);
  if (!got) return false;
  __threadfence_block();
  a = slot->a;
  b = slot->b;
  c = slot->c;
  n = slot->n;
  __syncwarp();
  if (lane == 0) *(volatile uint32_t*)&slot->ready = 0;  // slot may be refilled
  return true;
}

// =========================================================================
// k_send
// =========================================================================

constexpr uint32_t kFootCap = 2048;  // footers buffered per segment (ring offsets / 8)

struct SendPlanState {  // producer-only
  uint64_t rt, cap, staging, total_left, written_total, ncalls, cur, bidx;
  uint32_t partial, max_sge;
};

struct SendCallScratch {  // frames of the call being published
  const uint8_t* src[kMaxSgeLimit];
  uint64_t len[kMaxSgeLimit];
  uint64_t off[kMaxSgeLimit];
  uint32_t first_item[kMaxSgeLimit + 1];
};

// Producer: plan PairPollable::Send calls (pair.cc:645-734) one after another and publish
// their frames as work items until the op is finished or the footer buffer is full.
__device__ __noinline__ void send_produce_segment(const SendOpDev& op, const PairDev* P, SendPlanState& S,
                                                  SendCallScratch& CS, WorkItem* q, PipeCtl* ctl, uint32_t* foot8,
                                                  uint32_t* nfoot_out, uint32_t lane) {
  const uint64_t cap = S.cap, mask = cap - 1;
  uint32_t base_item = 0, nfoot = 0;
  bool op_done = false;
  while (nfoot + kMaxSgeLimit <= kFootCap) {
    const uint64_t rt = S.rt;
    const uint64_t rh = ld_acquire_u64(&P->credit_head);  // credit snapshot, once per call (pair.cc:650)
    const uint64_t cur = S.cur, bidx = S.bidx;
    const uint64_t idx = cur + lane;
    const bool valid = lane < S.max_sge && idx < op.nslices;
    const uint8_t* ptr = nullptr;
    uint64_t len = 0;
    if (valid) {
      SliceDev sl = op.slices[idx];
      const uint64_t skip = lane == 0 ? bidx : 0;
      ptr = sl.ptr + skip;
      len = sl.len - skip;
    }
    const uint64_t e = valid ? encoded_size(len) : 0;
    uint64_t incl = e;
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint64_t a = incl - e;  // staging / ring bytes consumed before this slice
    // min(CWS(send_buf_free), CWS(recv_buf_free)) (pair.cc:676-681): both shrink by `a`
    const uint64_t fr = free_size(cap, rh, rt);
    const uint64_t lim = S.staging < fr ? S.staging : fr;
    const uint64_t room = calc_writable(lim > a ? lim - a : 0);
    const bool fits = valid && len != 0 && len <= room;
    const unsigned bad = __ballot_sync(0xffffffffu, !fits);
    const int first_bad = __ffs(bad) - 1;
    const int nfull = first_bad < 0 ? 32 : first_bad;
    uint64_t p = 0;
    if ((int)lane < nfull) p = len;
    else if ((int)lane == nfull && valid && len != 0) p = room;  // cut: space ran out
    const unsigned fmask = __ballot_sync(0xffffffffu, p != 0);
    const uint32_t nframes = __popc(fmask);  // frames are lanes 0..nframes-1
    uint64_t wsum = p, esum = p ? encoded_size(p) : 0;
    const uint32_t items = p ? (uint32_t)((p + kChunk - 1) / kChunk) : 0;
    uint32_t items_incl = items;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, items_incl, o);
      if (lane >= (uint32_t)o) items_incl += t;
    }
    for (int o = 16; o > 0; o >>= 1) {
      wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
      esum += __shfl_xor_sync(0xffffffffu, esum, o);
    }
    const uint32_t nitems = __shfl_sync(0xffffffffu, items_incl, 31);
    const uint64_t cut_p = __shfl_sync(0xffffffffu, p, nfull < 32 ? nfull : 0);
    const uint64_t foff = (rt + a) & mask;
    CS.first_item[lane] = items_incl - items;
    if (p) {
      CS.src[lane] = ptr;
      CS.len[lane] = p;
      CS.off[lane] = foff;
      foot8[nfoot + lane] = (uint32_t)(((foff + 8 + round_up8(p)) & mask) >> 3);
    }
    if (lane == 0) {
      CS.first_item[32] = nitems;
      S.rt = (rt + esum) & mask;
      S.partial = wsum < S.total_left;  // pair.cc:712
      S.total_left -= wsum;
      S.written_total += wsum;
      if (wsum) S.ncalls++;
      // cursor advance (rdma_flush, rdma_bp_posix.cc:480-493)
      uint64_t nb = 0;
      if (nfull < 32 && nframes > (uint32_t)nfull) nb = (nfull == 0 ? bidx : 0) + cut_p;  // cut slice stays current
      else if (nfull == 0) nb = bidx;                                                    // nothing consumed
      S.cur = cur + nfull;
      S.bidx = nb;
    }
    __syncwarp();
    nfoot += nframes;
    // publish this call's items in id order, 32 at a time
    for (uint32_t it = lane; it < nitems; it += 32) {
      uint32_t f = 0;
      while (f + 1 < nframes && CS.first_item[f + 1] <= it) f++;
      const uint64_t c0 = (uint64_t)(it - CS.first_item[f]) * kChunk;
      const uint64_t flen = CS.len[f];
      uint64_t n = flen - c0;
      if (n > kChunk) n = kChunk;
      publish_item(q, base_item + it, reinterpret_cast<uint64_t>(CS.src[f] + c0), (CS.off[f] + 8 + c0) & mask,
                   c0 == 0 ? flen : 0, (uint32_t)n);
    }
    __syncwarp();
    base_item += nitems;
    const bool last = (wsum == 0) || !(op.flags & kFlagUntilBlocked) || S.total_left == 0;
    if (last) {
      op_done = true;
      break;
    }
  }
  if (lane == 0) {
    *nfoot_out = nfoot;
    ctl->total_items = base_item;
    ctl->op_done = op_done ? 1u : 0u;
    __threadfence_block();
    *(volatile uint32_t*)&ctl->seg_done = 1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 2)
k_send(PairDev* __restrict__ pairs, const SendOpDev* __restrict__ ops, OpResult* __restrict__ results) {
  __shared__ WorkItem q[kQI];
  __shared__ PipeCtl ctl;
  __shared__ SendPlanState PS;
  __shared__ SendCallScratch CS;
  __shared__ uint32_t foot8[kFootCap];
  __shared__ uint32_t s_nfoot;
  __shared__ unsigned long long s_total;
  __shared__ uint32_t s_status;
  const SendOpDev op = ops[blockIdx.x];
  PairDev* P = &pairs[op.slot];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid < kQI) q[tid].ready = 0;
  if (tid == 0) {
    s_total = 0;
    s_status = P->status;
    PS.rt = P->remote_tail;
    PS.cap = P->cap;
    PS.staging = P->cap / 2;  // send_buf_size = recv_buf_size / 2, pair.cc:104
    PS.max_sge = P->max_sge;
    PS.cur = 0;
    PS.bidx = op.byte_idx;
    PS.written_total = 0;
    PS.ncalls = 0;
    PS.partial = P->partial_write;
  }
  __syncthreads();
  {  // total_slice_size, pair.cc:661-664
    unsigned long long part = 0;
    for (uint64_t i = tid; i < op.nslices; i += kThreads) part += op.slices[i].len;
    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
    if (lane == 0 && part) atomicAdd(&s_total, part);
  }
  __syncthreads();
  if (s_status != kStConnected) {  // pair.cc:657
    if (tid == 0) {
      results[blockIdx.x].bytes = 0;
      results[blockIdx.x].calls = 0;
    }
    return;
  }
  if (tid == 0) PS.total_left = s_total - op.byte_idx;
  const uint64_t cap = P->cap, mask = cap - 1;
  uint8_t* ring = P->peer_ring;
  const bool sys_scope = P->wire != 0;

  while (true) {
    if (tid == 0) {
      ctl.next = 0;
      ctl.total_items = 0;
      ctl.seg_done = 0;
      ctl.op_done = 0;
      s_nfoot = 0;
    }
    __syncthreads();
    if (warp == 0) send_produce_segment(op, P, PS, CS, q, &ctl, foot8, &s_nfoot, lane);
    // ---------------------------------------------- move bytes
    uint64_t a, b, c;
    uint32_t n;
    while (claim_item(q, &ctl, lane, a, b, c, n)) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a);
      if (c != 0 && lane == 0) *reinterpret_cast<uint64_t*>(ring + ((b + cap - 8) & mask)) = c;  // AppendHeader
      uint64_t seg1 = cap - b;
      if (seg1 > n) seg1 = n;
      coop_copy<false>(ring + b, src, seg1, lane);
      if (n > seg1) coop_copy<false>(ring, src + seg1, n - seg1, lane);  // wrap: WR1 at remote+0
    }
    // footers last: a frame is complete for the reader only when header != 0 and footer == ~0
    // (ring_buffer.cc:75-96), so everything else of the segment is made visible first
    if (sys_scope) __threadfence_system();
    else __threadfence();
    __syncthreads();
    const uint32_t nfoot = s_nfoot;
    for (uint32_t i = tid; i < nfoot; i += kThreads)
      *reinterpret_cast<uint64_t*>(ring + ((uint64_t)foot8[i] << 3)) = kFooter;
    const bool done = ctl.op_done != 0;
    __syncthreads();
    if (done) break;
  }
  if (tid == 0) {
    P->remote_tail = PS.rt;
    P->partial_write = PS.partial;
    results[blockIdx.x].bytes = PS.written_total;
    results[blockIdx.x].calls = PS.ncalls;
    publish_mirror(P->mirror, P, P->mirror ? ((volatile PairMirror*)P->mirror)->has_message : 0,
                   P->mirror ? ((volatile PairMirror*)P->mirror)->readable : 0);
    // loopback wire: the peer lives in this table, refresh its readiness hint
    if (P->peer_slot >= 0 && PS.written_total) {
      __threadfence();
      PairDev* Q = &pairs[P->peer_slot];
      uint32_t hm;
      uint64_t rd;
      rx_probe(Q->ring, Q->cap, *(volatile uint64_t*)&Q->head, *(volatile uint64_t*)&Q->remain, hm, rd);
      if (Q->mirror) {
        volatile PairMirror* vm = Q->mirror;
        vm->has_message = hm;
        vm->readable = rd;
        __threadfence_system();
        vm->seq = vm->seq + 1;
      }
    }
  }
}

// =========================================================================
// k_recv
// =========================================================================
//
// The frames of a ring form a linked list (the next header sits right after the
// previous footer), so the producer is a scout: it walks the list through a
// 256-byte register window (one 8-byte word per lane: a 9-byte HTTP/2 header
// frame and the header of the payload frame behind it cost a single trip to
// memory), applies the Read/Recv integer logic and publishes 4 KiB items.
// Consumers: load, store to the destination slice, __syncwarp, then clear
// exactly the ring bytes just read (clear-on-read is part of the wire protocol,
// ring_buffer.cc:146,160,180).  A segment ends at a credit point or at the end.

struct ScoutState {  // producer-only, lives in shared memory between segments
  uint64_t head, mh, remain, acc, cap_left, delivered, ncalls;
  uint64_t credit_val;
  uint32_t credit_flag;
};

__device__ __forceinline__ uint64_t ld_volatile_u64(const void* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Producer: RingBufferPollable::Read (ring_buffer.cc:122-191) + PairPollable::Recv's credit
// rule (pair.cc:276-284) as integer logic over the frame list.  Two steps per batch of <= 32
// frames: (1) a minimal sequential walk of the list (header -> footer check -> next header)
// that leaves frame i in lane i; (2) everything else -- destination capacity, partial reads,
// pad/footer clearing, the credit threshold, work-item expansion -- lane-parallel with warp
// scans, exactly like the Send planner.
__device__ __noinline__ void recv_produce_segment(const RecvOpDev& op, const uint8_t* ring, uint64_t cap,
                                                  ScoutState& SS, WorkItem* q, PipeCtl* ctl, uint32_t lane) {
  const uint64_t mask = cap - 1;
  uint64_t head = SS.head, mh = SS.mh, remain = SS.remain, acc = SS.acc, cap_left = SS.cap_left;
  uint64_t delivered = SS.delivered, ncalls = SS.ncalls;
  uint64_t win = 0, win_base = 0;
  bool win_valid = false;
  uint32_t base_item = 0, credit = 0, last = 0;
  uint64_t credit_val = 0;
  uint64_t last_reload = 0;
  bool have_last = false;
  auto peek = [&](uint64_t off) -> uint64_t {  // 8-byte ring word at `off` through the window
    uint64_t d = (off - win_base) & mask;
    if (!win_valid || d >= 256) {
      // Frame lists are usually periodic (chttp2: 9-byte header frame + 16 KiB payload frame), so
      // the distance between the last two window reloads predicts where the next ones will be:
      // pull those lines into L2 now, 16 hops ahead, so the list walk is not one DRAM trip per hop.
      if (have_last) {
        const uint64_t stride = (off - last_reload) & mask;
        if (stride >= 256) {
          const uint64_t pf = (off + (uint64_t)((lane & 15) + 1) * stride + (lane >> 4) * 128) & mask;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ring + pf));
        }
      }
      last_reload = off;
      have_last = true;
      win_base = off;
      win = ld_volatile_u64(ring + ((off + 8ull * lane) & mask));
      win_valid = true;
      d = 0;
    }
    return __shfl_sync(0xffffffffu, win, (int)(d >> 3));
  };
  const bool one_call = !(op.flags & kFlagUntilBlocked);
  while (true) {
    // ---- step 1: walk the list; lane i keeps frame i
    uint64_t my_r = 0, my_head = 0;
    bool my_open = false;
    uint32_t cnt = 0;
    bool stopped = false;
    uint64_t h = head;
    if (remain > 0) {  // rest of a partially consumed frame (its header is already cleared)
      if (lane == 0) my_r = remain;
      cnt = 1;
    }
    const uint32_t want = one_call ? 1u : 32u;
    while (cnt < want) {  // GetReadableSize, ring_buffer.cc:67-97
      const uint64_t hdr = peek(h);
      if (hdr == 0 || hdr > cap - kReserved) { stopped = true; break; }
      const uint64_t foot = peek((h + 8 + round_up8(hdr)) & mask);
      if (foot != kFooter) { stopped = true; break; }
      if (lane == cnt) {
        my_r = hdr;
        my_head = h;
        my_open = true;
      }
      h = (h + 16 + round_up8(hdr)) & mask;
      cnt++;
    }
    // ---- step 2: Read()/Recv() per frame, all lanes at once
    const bool valid = lane < cnt;
    uint64_t r_incl = valid ? my_r : 0;
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, r_incl, o);
      if (lane >= (uint32_t)o) r_incl += t;
    }
    const uint64_t r_excl = r_incl - (valid ? my_r : 0);
    const uint64_t room = cap_left > r_excl ? cap_left - r_excl : 0;  // destination space left for this frame
    const uint64_t n = valid ? (my_r < room ? my_r : room) : 0;     // copy_size = min(readable, capacity)
    const bool full = valid && n == my_r && n != 0;
    const unsigned notfull = __ballot_sync(0xffffffffu, !full);
    const int first_nf = __ffs(notfull) - 1;
    uint32_t nproc = first_nf < 0 ? 32u : (uint32_t)first_nf;
    {  // a partially delivered frame is still processed (and is then the last one)
      const uint64_t n_at = __shfl_sync(0xffffffffu, n, nproc < 32 ? nproc : 0);
      if (nproc < 32 && n_at != 0) nproc++;
    }
    const uint64_t src = my_open ? (my_head + 8) & mask : mh;  // first payload byte to deliver
    const uint64_t end = (src + n) & mask;
    uint32_t ztail = 0;
    uint64_t mh_after = end;
    if (n == my_r) {  // frame finished: pad + footer, ring_buffer.cc:170-183
      const uint64_t up = round_up8(end);
      ztail = (uint32_t)(up - end) + 8;
      mh_after = ((up & mask) + 8) & mask;
    }
    const uint32_t zhead = my_open ? 8u : 0u;
    const bool proc = lane < nproc;
    // credit threshold (pair.cc:276-284): the first frame whose retired bytes push the
    // accumulator to cap/2 closes the segment
    uint64_t a_incl = proc ? (uint64_t)zhead + n + ztail : 0;  // internal_bytes_read of this call
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, a_incl, o);
      if (lane >= (uint32_t)o) a_incl += t;
    }
    const unsigned cross = __ballot_sync(0xffffffffu, proc && acc + a_incl >= cap / 2);
    if (cross) {
      const uint32_t ci = (uint32_t)__ffs(cross) - 1;
      nproc = ci + 1;
      credit = 1;
      credit_val = __shfl_sync(0xffffffffu, mh_after, ci);
    }
    if (nproc == 0) {  // nothing deliverable: empty ring, incomplete frame, or no room in dst
      last = 1;
      break;
    }
    const bool proc2 = lane < nproc;
    const uint32_t items = proc2 ? (uint32_t)((n + kChunk - 1) / kChunk) : 0;
    uint32_t items_incl = items;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, items_incl, o);
      if (lane >= (uint32_t)o) items_incl += t;
    }
    const uint32_t nitems = __shfl_sync(0xffffffffu, items_incl, 31);
    const uint32_t my_first = items_incl - items;
    // publish in id order: item `it` belongs to the frame f with first[f] <= it < first[f+1]
    for (uint32_t it0 = 0; it0 < nitems; it0 += 32) {
      const uint32_t it = it0 + lane;
      // find the owning frame by asking every lane whether it starts at or before `it`
      uint32_t f = 0;
      for (uint32_t g = 0; g < nproc; g++) {
        const uint32_t fg = __shfl_sync(0xffffffffu, my_first, g);
        const uint32_t ig = __shfl_sync(0xffffffffu, items, g);
        if (ig && fg <= it) f = g;
      }
      const uint64_t f_src = __shfl_sync(0xffffffffu, src, f);
      const uint64_t f_n = __shfl_sync(0xffffffffu, n, f);
      const uint64_t f_dst = delivered + __shfl_sync(0xffffffffu, r_excl, f);
      const uint32_t f_first = __shfl_sync(0xffffffffu, my_first, f);
      const uint32_t f_zh = __shfl_sync(0xffffffffu, zhead, f);
      const uint32_t f_zt = __shfl_sync(0xffffffffu, ztail, f);
      if (it < nitems) {
        const uint64_t c0 = (uint64_t)(it - f_first) * kChunk;
        uint64_t m = f_n - c0;
        const bool tail_item = m <= kChunk;
        if (m > kChunk) m = kChunk;
        const uint64_t z = (c0 == 0 ? f_zh : 0u) | ((uint64_t)(tail_item ? f_zt : 0u) << 16);
        publish_item(q, base_item + it, (f_src + c0) & mask, f_dst + c0, z, (uint32_t)m);
      }
    }
    __syncwarp();
    base_item += nitems;
    // ---- new cursor = state after the last processed frame
    const uint32_t L = nproc - 1;
    const bool open_L = __shfl_sync(0xffffffffu, (int)my_open, L) != 0;
    const uint64_t head_L = __shfl_sync(0xffffffffu, my_head, L);
    const uint64_t r_L = __shfl_sync(0xffffffffu, my_r, L);
    const uint64_t n_L = __shfl_sync(0xffffffffu, n, L);
    const uint64_t moved = __shfl_sync(0xffffffffu, r_excl, L) + n_L;
    if (open_L) head = (head_L + 16 + round_up8(r_L)) & mask;  // ring_buffer.cc:140-141
    mh = __shfl_sync(0xffffffffu, mh_after, L);
    remain = r_L - n_L;
    acc = credit ? 0 : acc + __shfl_sync(0xffffffffu, a_incl, L);
    delivered += moved;
    cap_left -= moved;
    ncalls += nproc;
    if (one_call || cap_left == 0 || (stopped && nproc == cnt)) last = 1;
    if (last || credit) break;
  }
  if (lane == 0) {
    SS.head = head;
    SS.mh = mh;
    SS.remain = remain;
    SS.acc = acc;
    SS.cap_left = cap_left;
    SS.delivered = delivered;
    SS.ncalls = ncalls;
    SS.credit_flag = credit;
    SS.credit_val = credit_val;
    ctl->total_items = base_item;
    ctl->op_done = last;
    __threadfence_block();
    *(volatile uint32_t*)&ctl->seg_done = 1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 2)
k_recv(PairDev* __restrict__ pairs, const RecvOpDev* __restrict__ ops, OpResult* __restrict__ results) {
  __shared__ WorkItem q[kQI];
  __shared__ PipeCtl ctl;
  __shared__ ScoutState SS;
  __shared__ uint32_t s_status;
  const RecvOpDev op = ops[blockIdx.x];
  PairDev* P = &pairs[op.slot];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid < kQI) q[tid].ready = 0;
  if (tid == 0) {
    s_status = P->status;
    SS.head = P->head;
    SS.mh = P->moving_head;
    SS.remain = P->remain;
    SS.acc = P->acc;
    SS.cap_left = op.cap;
    SS.delivered = 0;
    SS.ncalls = 0;
    SS.credit_flag = 0;
  }
  __syncthreads();
  if (s_status != kStConnected) {  // pair.cc:266-268
    if (tid == 0) {
      results[blockIdx.x].bytes = 0;
      results[blockIdx.x].calls = 0;
    }
    return;
  }
  uint8_t* ring = P->ring;
  const uint64_t cap = P->cap, mask = cap - 1;

  while (true) {
    if (tid == 0) {
      ctl.next = 0;
      ctl.total_items = 0;
      ctl.seg_done = 0;
      ctl.op_done = 0;
    }
    __syncthreads();
    if (warp == 0) recv_produce_segment(op, ring, cap, SS, q, &ctl, lane);
    uint64_t a, b, c;
    uint32_t n;
    auto clear_item = [&](uint64_t ia, uint64_t ic, uint32_t in) {  // clear-on-read: exactly what the item retired
      const uint32_t zhead = (uint32_t)(ic & 0xffff), ztail = (uint32_t)(ic >> 16);
      const uint64_t zs = (ia + cap - zhead) & mask;
      const uint64_t zl = (uint64_t)zhead + in + ztail;
      uint64_t z1 = cap - zs;
      if (z1 > zl) z1 = zl;
      coop_zero(ring + zs, z1, lane);
      if (zl > z1) coop_zero(ring, zl - z1, lane);
    };
    while (claim_item(q, &ctl, lane, a, b, c, n)) {
      // ---- scatter
      uint8_t* dst = op.dst + b;
      uint64_t seg1 = cap - a;
      if (seg1 > n) seg1 = n;
      coop_copy<true>(dst, ring + a, seg1, lane);
      if (n > seg1) coop_copy<true>(dst + seg1, ring, n - seg1, lane);
      __syncwarp();  // every lane's loads are done before any lane clears
      clear_item(a, c, n);
    }
    const bool credit = ld_shared_volatile(&SS.credit_flag) != 0;  // stable: the producer finished this segment
    if (credit) __threadfence_system();       // the sender may reuse the space only once it reads as zero
    __syncthreads();
    const bool done = ctl.op_done != 0;
    if (tid == 0 && credit) {
      // updateStatus, pair.cc:624-641: 16-byte status_report to the peer
      st_release_v2u64(P->peer_credit, SS.credit_val, 0);
      if (P->peer_mirror) ((volatile PairMirror*)P->peer_mirror)->credit_head = SS.credit_val;
      SS.credit_flag = 0;
    }
    __syncthreads();
    if (done) break;
  }
  if (tid == 0) {
    P->head = SS.head;
    P->moving_head = SS.mh;
    P->remain = SS.remain;
    P->acc = SS.acc;
    results[blockIdx.x].bytes = SS.delivered;
    results[blockIdx.x].calls = SS.ncalls;
    uint32_t hm;
    uint64_t rd;
    rx_probe(ring, cap, SS.head, SS.remain, hm, rd);
    publish_mirror(P->mirror, P, hm, rd);
  }
}

// =========================================================================
// k_poll_scan
// =========================================================================

__global__ void __launch_bounds__(128)
k_poll_scan(PairDev* __restrict__ pairs, const int32_t* __restrict__ slots, uint32_t* __restrict__ events,
            uint32_t* __restrict__ ready_count, int32_t* __restrict__ ready_slots, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t ev = 0;
  int32_t slot = -1;
  if (i < n) {
    slot = slots[i];
    PairDev* P = &pairs[slot];
    const uint32_t st = *(volatile uint32_t*)&P->status;
    if (st == kStConnected) {
      const uint32_t exit_flag = ld_acquire_u32(&P->credit_exit);
      uint32_t hm;
      uint64_t rd;
      rx_probe(P->ring, P->cap, *(volatile uint64_t*)&P->head, *(volatile uint64_t*)&P->remain, hm, rd);
      const uint32_t pw = *(volatile uint32_t*)&P->partial_write;
      if (exit_flag == 1) {
        ev = kEvReadable;  // HalfClosed: force a read event (engine :1130-1137)
      } else {
        if (hm) ev |= kEvReadable;
        if (pw) ev |= kEvWritable;
      }
      publish_mirror(P->mirror, P, hm, rd);
    } else if (st == kStError || st == kStHalfClosed) {
      ev = kEvReadable;
    }
    events[i] = ev;
  }
  // warp-aggregated append to the ready set
  const unsigned m = __ballot_sync(0xffffffffu, ev != 0);
  if (m) {
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(ready_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (ev) ready_slots[base + __popc(m & ((1u << lane) - 1))] = slot;
  }
}

// =========================================================================
// k_probe_copy: calibration kernel.  Same decomposition as k_send (one CTA per
// connection, 4 KiB warp items, same copy primitives) but no framing logic: what
// this grid shape can reach on this GPU, for src/dst misalignments `mis`.
// =========================================================================
__global__ void __launch_bounds__(512, 2)
k_probe_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint64_t bytes_per_cta, uint64_t stride,
             uint32_t mis, uint32_t item_bytes, uint32_t dynamic) {
  __shared__ uint32_t s_next;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint8_t* d = dst + (uint64_t)blockIdx.x * stride;
  const uint8_t* sp = src + (uint64_t)blockIdx.x * stride + mis;
  const uint64_t nitems = bytes_per_cta / item_bytes;
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  const bool zero_after = (dynamic & 2) != 0;  // recv-like traffic: read src, write dst, clear src
  uint8_t* sw = const_cast<uint8_t*>(sp);
  if (dynamic & 1) {
    while (true) {
      uint32_t w = 0;
      if (lane == 0) w = atomicAdd(&s_next, 1u);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= nitems) break;
      if (zero_after) {
        coop_copy<true>(d + (uint64_t)w * item_bytes, sp + (uint64_t)w * item_bytes, item_bytes, lane);
        __syncwarp();
        coop_zero(sw + (uint64_t)w * item_bytes, item_bytes, lane);
      } else {
        coop_copy<false>(d + (uint64_t)w * item_bytes, sp + (uint64_t)w * item_bytes, item_bytes, lane);
      }
    }
  } else {
    for (uint64_t w = warp; w < nitems; w += nwarps) {
      if (zero_after) {
        coop_copy<true>(d + w * item_bytes, sp + w * item_bytes, item_bytes, lane);
        __syncwarp();
        coop_zero(sw + w * item_bytes, item_bytes, lane);
      } else {
        coop_copy<false>(d + w * item_bytes, sp + w * item_bytes, item_bytes, lane);
      }
    }
  }
}

void launch_probe_copy(uint8_t* dst, const uint8_t* src, uint64_t bytes_per_cta, uint64_t stride, int nctas,
                       int threads, uint32_t mis, uint32_t item_bytes, uint32_t dynamic, void* stream) {
  k_probe_copy<<<nctas, threads, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, bytes_per_cta, stride, mis,
                                                                         item_bytes, dynamic);
}

// ---------------------------------------------------------------- launchers

void launch_send(PairDev* pairs, const SendOpDev* ops, OpResult* results, int nops, void* stream) {
  if (nops <= 0) return;
  k_send<<<nops, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(pairs, ops, results);
}
void launch_recv(PairDev* pairs, const RecvOpDev* ops, OpResult* results, int nops, void* stream) {
  if (nops <= 0) return;
  k_recv<<<nops, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(pairs, ops, results);
}
void launch_poll_scan(PairDev* pairs, const int32_t* slots, uint32_t* events, uint32_t* ready_count,
                      int32_t* ready_slots, int n, void* stream) {
  if (n <= 0) return;
  k_poll_scan<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(pairs, slots, events, ready_count,
                                                                               ready_slots, n);
}

}  // namespace b200
