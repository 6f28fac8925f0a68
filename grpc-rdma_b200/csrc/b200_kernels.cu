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
template <bool kRingSrc, bool kHi, bool kShift>
__device__ __noinline__ void copy_vec_shifted(uint8_t* dst, const uint8_t* src_al, uint64_t nvec, unsigned sh,
                                              uint32_t lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src_al);
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint64_t step = 32ull * kUnroll;
  const uint64_t nfull = nvec / step * step;
  if (nfull) {
    uint4 carry = ld_src16<kRingSrc>(s + lane);
    for (uint64_t base = 0; base < nfull; base += step) {
      uint4 A[kUnroll + 1];
      A[0] = carry;
#pragma unroll
      for (int k = 1; k < kUnroll; k++) A[k] = ld_src16<kRingSrc>(s + base + 32ull * k + lane);
      {
        const uint64_t i = base + step + lane;  // next block's first row; only words <= nvec exist
        A[kUnroll] = i <= nvec ? ld_src16<kRingSrc>(s + i) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < kUnroll; k++) {
        const uint4 B = shfl_down1_wrap(A[k], A[k + 1], lane);
        st_stream16(d + base + 32ull * k + lane, shift_window<kHi, kShift>(A[k], B, sh));
      }
      carry = A[kUnroll];
    }
  }
  for (uint64_t i = nfull + lane; i < nvec; i += 32) {
    uint4 A = ld_src16<kRingSrc>(s + i), B = ld_src16<kRingSrc>(s + i + 1);
    st_stream16(d + i, shift_window<kHi, kShift>(A, B, sh));
  }
}

template <bool kRingSrc>
__device__ __noinline__ void copy_vec_aligned(uint8_t* dst, const uint8_t* src, uint64_t nvec, uint32_t lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint64_t step = 32ull * kUnroll;
  const uint64_t nfull = nvec / step * step;
  for (uint64_t base = lane; base < nfull; base += step) {
    uint4 v[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; k++) v[k] = ld_src16<kRingSrc>(s + base + 32ull * k);
#pragma unroll
    for (int k = 0; k < kUnroll; k++) st_stream16(d + base + 32ull * k, v[k]);
  }
  for (uint64_t i = nfull + lane; i < nvec; i += 32) st_stream16(d + i, ld_src16<kRingSrc>(s + i));
}

// Warp-cooperative copy of n bytes, any alignment on either side.  Bulk = aligned 16-byte
// stores; the <16-byte head and tail use one byte per lane.
template <bool kRingSrc>
__device__ __forceinline__ void coop_copy(uint8_t* dst, const uint8_t* src, uint64_t n, uint32_t lane) {
  if (n == 0) return;
  uint64_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head > n) head = n;
  if (lane < head) dst[lane] = kRingSrc ? *reinterpret_cast<const volatile uint8_t*>(src + lane) : src[lane];
  dst += head;
  src += head;
  n -= head;
  const uint64_t nvec = n >> 4;
  const unsigned m = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15);
  if (nvec) {
    if (m == 0) {
      copy_vec_aligned<kRingSrc>(dst, src, nvec, lane);
    } else {
      const uint8_t* sal = src - m;
      const unsigned sh = (m & 7) * 8;
      if (m & 8) {
        if (sh) copy_vec_shifted<kRingSrc, true, true>(dst, sal, nvec, sh, lane);
        else copy_vec_shifted<kRingSrc, true, false>(dst, sal, nvec, sh, lane);
      } else {
        copy_vec_shifted<kRingSrc, false, true>(dst, sal, nvec, sh, lane);
      }
    }
  }
  const uint64_t done = nvec << 4;
  const uint64_t tail = n - done;
  if (lane < tail)
    dst[done + lane] = kRingSrc ? *reinterpret_cast<const volatile uint8_t*>(src + done + lane) : src[done + lane];
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
// k_send
// =========================================================================
//
// One CTA per (pair, write).  Warp 0 plans Send() call j+1 (credit snapshot,
// warp scan of encoded sizes, cut detection) while all warps -- warp 0 joins
// once the plan is published -- move the bytes of call j in 4 KiB work items
// claimed from a shared counter.  One CTA barrier per call; footers of call j
// are written after that barrier (everything else of the call is fenced before).

constexpr int kSendThreads = 512;
constexpr uint32_t kChunk = 4096;  // payload bytes per work item

struct FrameDesc {
  const uint8_t* src;
  uint64_t len;   // payload bytes
  uint64_t off;   // ring offset of the frame header
  uint32_t first_item;
  uint32_t _pad;
};

struct SendCall {
  FrameDesc frames[kMaxSgeLimit];
  uint32_t nframes, nitems, last, _pad;
};

struct SendPlanState {  // touched by warp 0 only
  uint64_t rt, cap, staging, total_left, written_total, ncalls, cur, bidx;
  uint32_t partial, max_sge;
};

// One PairPollable::Send call (pair.cc:645-734) as integer planning; executed by
// warp 0, all lanes converged.  Returns true when this was the last call.
__device__ __forceinline__ void plan_send_call(const SendOpDev& op, const PairDev* P, SendPlanState& S,
                                               SendCall& out, uint32_t lane) {
  const uint64_t cap = S.cap, mask = cap - 1, rt = S.rt;
  const uint64_t rh = ld_acquire_u64(&P->credit_head);  // credit snapshot, once per call (pair.cc:650)
  const uint64_t cur = S.cur, bidx = S.bidx;
  const uint64_t idx = cur + lane;
  const bool valid = lane < S.max_sge && idx < op.nslices;
  const uint8_t* ptr = nullptr;
  uint64_t len = 0;
  if (valid) {
    SliceDev sl = op.slices[idx];
    uint64_t skip = lane == 0 ? bidx : 0;
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
  const uint32_t nframes = __popc(fmask);
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
  if (p) {
    FrameDesc& f = out.frames[lane];  // frames are lanes 0..nframes-1
    f.src = ptr;
    f.len = p;
    f.off = (rt + a) & mask;
    f.first_item = items_incl - items;
  }
  const uint32_t nitems = __shfl_sync(0xffffffffu, items_incl, 31);
  const uint64_t cut_p = __shfl_sync(0xffffffffu, p, nfull < 32 ? nfull : 0);
  if (lane == 0) {
    out.nframes = nframes;
    out.nitems = nitems;
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
    out.last = (wsum == 0) || !(op.flags & kFlagUntilBlocked) || S.total_left == 0;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kSendThreads, 2)
k_send(PairDev* __restrict__ pairs, const SendOpDev* __restrict__ ops, OpResult* __restrict__ results) {
  __shared__ SendCall calls[2];
  __shared__ SendPlanState PS;
  __shared__ unsigned long long s_total;
  __shared__ uint32_t s_next[2];
  __shared__ uint32_t s_status;
  const SendOpDev op = ops[blockIdx.x];
  PairDev* P = &pairs[op.slot];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    s_total = 0;
    s_next[0] = s_next[1] = 0;
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
    for (uint64_t i = tid; i < op.nslices; i += kSendThreads) part += op.slices[i].len;
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
  __syncthreads();
  const uint64_t cap = PS.cap, mask = cap - 1;
  uint8_t* ring = P->peer_ring;
  const bool sys_scope = P->wire != 0;

  if (warp == 0) plan_send_call(op, P, PS, calls[0], lane);
  __syncthreads();

  for (uint32_t j = 0;; j++) {
    SendCall& cur = calls[j & 1];
    const bool last = cur.last != 0;
    // footers of the previous call: everything else of it was fenced before the barrier.
    // A frame is complete for the reader only when header != 0 and footer == ~0
    // (ring_buffer.cc:75-96), so the footer goes last.
    if (j > 0) {
      const SendCall& prev = calls[(j - 1) & 1];
      if (tid >= 32 && tid - 32 < prev.nframes) {
        const FrameDesc fd = prev.frames[tid - 32];
        *reinterpret_cast<uint64_t*>(ring + ((fd.off + 8 + round_up8(fd.len)) & mask)) = kFooter;
      }
    }
    __syncthreads();  // prev.frames fully consumed before warp 0 overwrites that buffer
    if (warp == 0 && !last) plan_send_call(op, P, PS, calls[(j + 1) & 1], lane);
    // ---------------------------------------------- move the bytes of call j
    const uint32_t nframes = cur.nframes, nitems = cur.nitems;
    while (true) {
      uint32_t w = 0;
      if (lane == 0) w = atomicAdd(&s_next[j & 1], 1u);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= nitems) break;
      uint32_t f = 0;
      while (f + 1 < nframes && cur.frames[f + 1].first_item <= w) f++;
      const FrameDesc fd = cur.frames[f];
      const uint64_t c0 = (uint64_t)(w - fd.first_item) * kChunk;
      uint64_t n = fd.len - c0;
      if (n > kChunk) n = kChunk;
      if (c0 == 0 && lane == 0) *reinterpret_cast<uint64_t*>(ring + fd.off) = fd.len;  // AppendHeader
      const uint64_t pos = (fd.off + 8 + c0) & mask;
      uint64_t seg1 = cap - pos;
      if (seg1 > n) seg1 = n;
      coop_copy<false>(ring + pos, fd.src + c0, seg1, lane);
      if (n > seg1) coop_copy<false>(ring, fd.src + c0 + seg1, n - seg1, lane);  // wrap: WR1 at remote+0
    }
    if (sys_scope) __threadfence_system();
    else __threadfence();
    __syncthreads();
    if (tid == 0) s_next[j & 1] = 0;
    if (last) {
      if (tid < nframes) {
        const FrameDesc fd = cur.frames[tid];
        *reinterpret_cast<uint64_t*>(ring + ((fd.off + 8 + round_up8(fd.len)) & mask)) = kFooter;
      }
      break;
    }
  }
  __syncthreads();
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
// One CTA per (pair, read).  The frames of a ring form a linked list (the next
// header sits right after the previous footer), so warp 0 is a scout: it walks
// the list through a 256-byte register window (one 8-byte word per lane, so a
// 9-byte HTTP/2 header frame and the header of the payload frame behind it cost
// a single trip to memory), applies the Read/Recv integer logic and queues
// batches of frames.  While the scout walks batch b+1, all warps scatter batch b
// in 4 KiB work items: load, store to the destination slice, __syncwarp, then
// clear exactly the ring bytes this warp just read (clear-on-read is part of
// the wire protocol, ring_buffer.cc:146,160,180).  One CTA barrier per batch.

constexpr int kRecvThreads = 512;
constexpr int kBatchFrames = 32;
constexpr uint64_t kBatchBytes = 192 * 1024;  // close a batch once this much payload is queued

struct RecvFrame {
  uint64_t src_off;   // ring offset of the first payload byte to deliver
  uint64_t n;         // bytes to deliver
  uint64_t dst_off;   // offset in the destination
  uint32_t zhead;     // bytes to clear before the payload (the header word on first touch)
  uint32_t ztail;     // bytes to clear after it (pad + footer once the frame is finished)
  uint32_t first_item;
  uint32_t _pad;
};

struct RecvBatch {
  RecvFrame f[kBatchFrames];
  uint64_t credit_val;
  uint32_t nframes, nitems, credit_flag, last;
};

struct ScoutState {  // registers of warp 0, uniform across lanes
  uint64_t head, mh, remain, acc, cap_left, delivered, ncalls;
  uint64_t win, win_base;
  bool win_valid;
};

__device__ __forceinline__ uint64_t ld_volatile_u64(const void* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// 8-byte word at ring offset `off` through the scout's window
__device__ __forceinline__ uint64_t scout_peek(ScoutState& st, const uint8_t* ring, uint64_t mask, uint64_t off,
                                               uint32_t lane) {
  uint64_t d = (off - st.win_base) & mask;
  if (!st.win_valid || d >= 256) {
    st.win_base = off;
    st.win = ld_volatile_u64(ring + ((off + 8ull * lane) & mask));
    st.win_valid = true;
    d = 0;
  }
  return __shfl_sync(0xffffffffu, st.win, (int)(d >> 3));
}

// Queue the next batch: RingBufferPollable::Read (ring_buffer.cc:122-191) + PairPollable::Recv's
// credit rule (pair.cc:276-284) as integer logic over the frame list.
__device__ __forceinline__ void scout_batch(ScoutState& st, const uint8_t* ring, uint64_t cap, const RecvOpDev& op,
                                            RecvBatch& out, uint32_t lane) {
  const uint64_t mask = cap - 1;
  uint32_t nframes = 0, nitems = 0, credit = 0, last = 0;
  uint64_t bytes = 0, credit_val = 0;
  while (nframes < (uint32_t)kBatchFrames && bytes < kBatchBytes) {
    uint64_t r;
    bool opening = false;
    if (st.remain > 0) {
      r = st.remain;
    } else {  // GetReadableSize, ring_buffer.cc:67-97
      const uint64_t hdr = scout_peek(st, ring, mask, st.head, lane);
      if (hdr == 0 || hdr > cap - kReserved) { last = 1; break; }
      const uint64_t foot = scout_peek(st, ring, mask, (st.head + 8 + round_up8(hdr)) & mask, lane);
      if (foot != kFooter) { last = 1; break; }
      r = hdr;
      opening = true;
    }
    const uint64_t n = r < st.cap_left ? r : st.cap_left;
    if (n == 0) { last = 1; break; }
    if (opening) {  // first touch of this frame, ring_buffer.cc:135-147
      st.mh = (st.head + 8) & mask;
      st.head = (st.head + 16 + round_up8(r)) & mask;
    }
    const uint64_t src_off = st.mh;
    st.mh = (st.mh + n) & mask;
    st.remain = r - n;
    uint32_t ztail = 0;
    if (st.remain == 0) {  // pad + footer, ring_buffer.cc:170-183
      const uint64_t up = round_up8(st.mh);
      ztail = (uint32_t)(up - st.mh) + 8;
      st.mh = ((up & mask) + 8) & mask;
    }
    const uint32_t zhead = opening ? 8u : 0u;
    const uint32_t items = (uint32_t)((n + kChunk - 1) / kChunk);
    if (lane == 0) {
      RecvFrame& f = out.f[nframes];
      f.src_off = src_off;
      f.n = n;
      f.dst_off = st.delivered;
      f.zhead = zhead;
      f.ztail = ztail;
      f.first_item = nitems;
    }
    nframes++;
    nitems += items;
    bytes += n;
    st.delivered += n;
    st.cap_left -= n;
    st.ncalls++;
    st.acc += (uint64_t)zhead + n + ztail;  // internal_bytes_read of this call
    if (st.acc >= cap / 2) {                // pair.cc:276-284: credit goes out after this batch is cleared
      credit = 1;
      credit_val = st.mh;
      st.acc = 0;
    }
    if (!(op.flags & kFlagUntilBlocked) || st.cap_left == 0) { last = 1; break; }
    if (credit) break;
  }
  if (lane == 0) {
    out.nframes = nframes;
    out.nitems = nitems;
    out.credit_flag = credit;
    out.credit_val = credit_val;
    out.last = last;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kRecvThreads, 2)
k_recv(PairDev* __restrict__ pairs, const RecvOpDev* __restrict__ ops, OpResult* __restrict__ results) {
  __shared__ RecvBatch batches[2];
  __shared__ uint32_t s_next[2];
  __shared__ uint32_t s_status;
  const RecvOpDev op = ops[blockIdx.x];
  PairDev* P = &pairs[op.slot];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    s_status = P->status;
    s_next[0] = s_next[1] = 0;
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

  ScoutState st;
  if (warp == 0) {
    st.head = P->head;
    st.mh = P->moving_head;
    st.remain = P->remain;
    st.acc = P->acc;
    st.cap_left = op.cap;
    st.delivered = 0;
    st.ncalls = 0;
    st.win = 0;
    st.win_base = 0;
    st.win_valid = false;
    scout_batch(st, ring, cap, op, batches[0], lane);
  }
  __syncthreads();

  for (uint32_t b = 0;; b++) {
    RecvBatch& cur = batches[b & 1];
    const bool last = cur.last != 0;
    if (warp == 0 && !last) scout_batch(st, ring, cap, op, batches[(b + 1) & 1], lane);
    const uint32_t nframes = cur.nframes, nitems = cur.nitems;
    while (true) {
      uint32_t w = 0;
      if (lane == 0) w = atomicAdd(&s_next[b & 1], 1u);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= nitems) break;
      uint32_t fi = 0;
      while (fi + 1 < nframes && cur.f[fi + 1].first_item <= w) fi++;
      const RecvFrame fr = cur.f[fi];
      const uint64_t c0 = (uint64_t)(w - fr.first_item) * kChunk;
      uint64_t n = fr.n - c0;
      const bool tail_item = n <= kChunk;
      if (n > kChunk) n = kChunk;
      // ---- scatter
      const uint64_t pos = (fr.src_off + c0) & mask;
      uint8_t* dst = op.dst + fr.dst_off + c0;
      uint64_t seg1 = cap - pos;
      if (seg1 > n) seg1 = n;
      coop_copy<true>(dst, ring + pos, seg1, lane);
      if (n > seg1) coop_copy<true>(dst + seg1, ring, n - seg1, lane);
      __syncwarp();  // every lane's loads are done before any lane clears
      // ---- clear-on-read: exactly what this item retired
      uint64_t zs = pos, zl = n;
      if (c0 == 0) {
        zs = (pos + cap - fr.zhead) & mask;
        zl += fr.zhead;
      }
      if (tail_item) zl += fr.ztail;
      uint64_t z1 = cap - zs;
      if (z1 > zl) z1 = zl;
      coop_zero(ring + zs, z1, lane);
      if (zl > z1) coop_zero(ring, zl - z1, lane);
    }
    const bool credit = cur.credit_flag != 0;
    if (credit) __threadfence_system();  // the sender may reuse the space only once it reads as zero
    __syncthreads();
    if (tid == 0) {
      s_next[b & 1] = 0;
      if (credit) {
        // updateStatus, pair.cc:624-641: 16-byte status_report to the peer
        st_release_v2u64(P->peer_credit, cur.credit_val, 0);
        if (P->peer_mirror) ((volatile PairMirror*)P->peer_mirror)->credit_head = cur.credit_val;
      }
    }
    if (last) break;
    // nobody may still be reading cur (= the buffer the scout fills next iteration) -- guaranteed
    // by the barrier above; the scout's writes to the other buffer finished before it as well
  }
  if (tid == 0) {  // warp 0 lane 0 holds the final cursor
    P->head = st.head;
    P->moving_head = st.mh;
    P->remain = st.remain;
    P->acc = st.acc;
    results[blockIdx.x].bytes = st.delivered;
    results[blockIdx.x].calls = st.ncalls;
    uint32_t hm;
    uint64_t rd;
    rx_probe(ring, cap, st.head, st.remain, hm, rd);
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

// ---------------------------------------------------------------- launchers

void launch_send(PairDev* pairs, const SendOpDev* ops, OpResult* results, int nops, void* stream) {
  if (nops <= 0) return;
  k_send<<<nops, kSendThreads, 0, static_cast<cudaStream_t>(stream)>>>(pairs, ops, results);
}
void launch_recv(PairDev* pairs, const RecvOpDev* ops, OpResult* results, int nops, void* stream) {
  if (nops <= 0) return;
  k_recv<<<nops, kRecvThreads, 0, static_cast<cudaStream_t>(stream)>>>(pairs, ops, results);
}
void launch_poll_scan(PairDev* pairs, const int32_t* slots, uint32_t* events, uint32_t* ready_count,
                      int32_t* ready_slots, int n, void* stream) {
  if (n <= 0) return;
  k_poll_scan<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(pairs, slots, events, ready_count,
                                                                               ready_slots, n);
}

}  // namespace b200
