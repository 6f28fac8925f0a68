// b200_dev.cuh -- device-visible state of a connection and the integer helpers
// shared by the sm_100a kernels and the host runtime.
//
// Reference being replaced (paths relative to the reference root):
//   RingBufferPollable state      src/core/lib/ibverbs/ring_buffer.h:203-208
//   PairPollable cursors/credit   src/core/lib/ibverbs/pair.h:100-103,168-172
//   credit / framing arithmetic   src/core/lib/ibverbs/ring_buffer.h:180-189,
//                                 ring_buffer.cc:99-116
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

constexpr uint64_t kAlign = 8;          // ring_buffer.h:49
constexpr uint64_t kReserved = 24;      // ring_buffer.h:52 (header, footer, one spare word)
constexpr uint64_t kFooter = ~0ull;     // ring_buffer.h:50
constexpr int kMaxSgeLimit = 32;        // planner handles one warp of slices per Send call

B200_HD uint64_t round_up8(uint64_t v) { return (v + 7) & ~7ull; }
B200_HD uint64_t round_down8(uint64_t v) { return v & ~7ull; }
// GetEncodedSize, ring_buffer.h:180-183
B200_HD uint64_t encoded_size(uint64_t payload) { return 16 + round_up8(payload); }
// CalculateWritableSize, ring_buffer.h:185-189
B200_HD uint64_t calc_writable(uint64_t space) { return space > kReserved ? round_down8(space - kReserved) : 0; }
// GetFreeSize, ring_buffer.cc:99-104
B200_HD uint64_t free_size(uint64_t cap, uint64_t head, uint64_t tail) {
  return cap - ((tail + cap - head) & (cap - 1));
}
// GetWritableSize(head, tail), ring_buffer.cc:106-116
B200_HD uint64_t writable_size(uint64_t cap, uint64_t head, uint64_t tail) {
  uint64_t f = free_size(cap, head, tail);
  return f > kReserved ? f - kReserved : 0;
}

// Host-visible mirror of one pair (pinned, GPU-mapped).  Kernels refresh the
// fields they own at the end of every op that touches the pair (posted writes
// only) so that HasMessage / HasPendingWrites / get_status stay wait-free host
// reads (pair.cc:288-303).
struct PairMirror {
  uint64_t head, moving_head, remain, acc;
  uint64_t remote_tail, credit_head;
  uint64_t readable;       // GetReadableSize() as of the last refresh
  uint32_t partial_write;  // HasPendingWrites()
  uint32_t peer_exit;      // status_report.peer_exit seen by this pair
  uint32_t has_message;    // HasMessage()
  uint32_t _reserved;
};

// One connection endpoint in HBM.  Three blocks with distinct writers:
//   setup  : written by the host at Init/Connect/Disconnect only
//   cursor : owned by this pair's own send/recv kernels
//   credit : the 16-byte status_report the PEER writes (pair.h:100-103);
//            16-byte aligned so one v2.u64 store updates it atomically
struct __align__(128) PairDev {
  // ---- setup
  uint8_t* ring;         // this pair's receive ring (HBM), all-zero when empty
  uint64_t cap;          // power of two (ring_buffer.cc:22)
  uint8_t* peer_ring;    // where Send lands frames: the peer's ring (remote_addr)
  uint64_t* peer_credit; // address of the peer's credit block
  PairMirror* mirror;
  PairMirror* peer_mirror;  // loopback wire only, else nullptr
  uint32_t status;       // b200_status as the host last set it
  uint32_t max_sge;      // frames per Send call (pair.cc:672)
  int32_t peer_slot;     // loopback wire: index of the peer in this table, else -1
  uint32_t wire;         // 0 = same device, 1 = peer device over NVLink (system-scope fences)
  // ---- cursor
  uint64_t head;         // RingBufferPollable::head_
  uint64_t moving_head;  // moving_head_
  uint64_t remain;       // remain_
  uint64_t acc;          // PairPollable::internal_read_size_
  uint64_t remote_tail;  // PairPollable::remote_tail_
  uint32_t partial_write;
  uint32_t mlock;        // serialises "read the device truth, write the mirror" when ops of the two
                         // ends run concurrently (kFlagConcurrent)
  // ---- credit (offset 112, 16-byte aligned)
  uint64_t credit_head;  // status_report.remote_head
  uint32_t credit_exit;  // status_report.peer_exit
  uint32_t _pad1;        // covered by the peer's 16-byte status write: nothing of ours may live here
};
static_assert(sizeof(PairDev) == 128, "PairDev is one 128-byte line");

struct SliceDev {  // same layout as b200_slice
  const uint8_t* ptr;
  uint64_t len;
};

struct SendOpDev {
  int32_t slot;
  uint32_t flags;  // B200_BATCH_*
  const SliceDev* slices;
  uint64_t nslices;
  uint64_t byte_idx;
  uint64_t nreal;  // slices [nreal, nslices) only count towards total_slice_size (pair.cc:661-664): the host folds
                   // what lies beyond max_sge into one trailing pseudo-slice that must never be dereferenced
};

struct RecvOpDev {
  int32_t slot;
  uint32_t flags;
  uint8_t* dst;
  uint64_t cap;
};

// per-op result: bytes moved and number of Send/Recv calls that moved > 0
struct OpResult {
  uint64_t bytes;
  uint64_t calls;
};

constexpr uint32_t kFlagUntilBlocked = 0x1;
constexpr uint32_t kFlagConcurrent = 0x8;  // B200_BATCH_CONCURRENT
constexpr uint32_t kEvReadable = 0x1;
constexpr uint32_t kEvWritable = 0x4;

constexpr uint32_t kStConnected = 2;
constexpr uint32_t kStHalfClosed = 3;
constexpr uint32_t kStError = 5;

// ---- persistent service (b200_service_*): three resident kernels.
//   owners  (k_svc_owner): one WARP per command queue.  The host posts 128-byte commands into the
//           queue's ring in pinned mapped memory; the warp polls it (two entries per trip over PCIe),
//           executes small Send / Recv calls entirely by itself -- plan, copy, cursors, credit,
//           mirrors: no CTA barrier, no lock (both ends of a loopback connection map to the same
//           owner, so everything that touches a connection's small ops is program-ordered) -- and
//           hands anything larger to the pool through a mailbox in device memory.
//   pool    (k_svc_big): CTAs that run send_body / recv_body (the k_send / k_recv code) on mailbox
//           jobs and answer the host themselves.
//   poller  (k_svc_poll): the resident readiness scan of the BPEV design (ready ring).
constexpr uint32_t kSvcSend = 1, kSvcRecv = 2, kSvcStop = 3, kSvcRetire = 4, kSvcNop = 5;
constexpr uint32_t kSvcInline = 5;     // slices carried inside the command (a unary call has 2-4)
constexpr uint32_t kOwnQ = 16;         // command entries per owner queue
constexpr uint32_t kOwnBoxes = 8;      // pool jobs in flight per owner
constexpr uint32_t kSmallMax = 8192;   // bytes one warp moves by itself; larger ops go to the pool
constexpr uint32_t kEagerMax = 2048;   // frames up to this size are pushed to the receiver's host slot
constexpr uint32_t kSvcSliceArea = 1024;  // pinned slice descriptors per command entry
struct __align__(128) SvcCmd {
  uint32_t stamp;    // ticket + 1, written last by the host (first 64-byte half of the line)
  uint32_t op;       // kSvc*
  int32_t slot;
  uint32_t flags;    // B200_BATCH_*
  uint64_t ptr;      // send: SliceDev* (GPU-addressable; unused when n <= kSvcInline)   recv: destination
  uint64_t n;        // send: nslices                        recv / retire: capacity
  uint64_t byte_idx;
  SliceDev inl[kSvcInline];  // send: the slice list itself when it is short -- no second trip over PCIe
  uint32_t nreal;    // send: slices [nreal, n) only count towards total_slice_size (never dereferenced)
  uint32_t stamp2;   // = stamp, in the second 64-byte half: the two halves may be read by separate PCIe reads
};
static_assert(sizeof(SvcCmd) == 128, "one command = one 128-byte line");
struct __align__(16) SvcDone {  // one 16-byte store: the fields become visible together
  uint64_t bytes;
  uint32_t calls;
  uint32_t seq;      // = SvcCmd.stamp once the op is finished and its bytes are visible
};
static_assert(sizeof(SvcDone) == 16, "one answer = one 16-byte store");
// service-only device state of a pair
struct PairSvc {
  uint64_t delivered;  // payload bytes this pair's Recv calls have returned since the service started
  uint64_t pushed_at;  // value of `delivered` for which the frame at the head was pushed to the host slot (~0: none)
};
// host-visible record of an eagerly pushed frame (pinned): the frame at the head of the ring, complete and
// <= kEagerMax bytes, copied to the pair's host slot so that Recv does not need a trip to the GPU.  No
// ordering between the payload stores and the record is assumed: `csum` covers payload, size and `at`.
struct __align__(32) EagerRec {
  uint64_t at;     // = PairSvc.delivered when the frame was pushed: valid only while the host's count agrees
  uint64_t csum;
  uint32_t size;
  uint32_t magic;
  uint64_t _pad;
};
constexpr uint32_t kEagerMagic = 0xEA6E7001u;
// mailbox owner -> pool (device memory)
struct __align__(128) BigBox {
  uint32_t state;    // 0 free, 1 posted, 2 running, 3 done (owner reaps)
  uint32_t kind;     // kSvcSend / kSvcRecv
  int32_t slot;
  uint32_t flags;
  uint64_t ptr, n, byte_idx, nreal;
  SliceDev inl[kSvcInline];
  SvcDone* done;     // host entry the pool answers into
  uint32_t seq;
  uint32_t _pad;
  OpResult res;
};
B200_HD uint64_t eager_mix(uint64_t x) {
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  return x;
}
// checksum contribution of payload word j (tail bytes beyond `size` zeroed); XOR of all + eager_mix(at * 31 + size)
B200_HD uint64_t eager_word(uint64_t w, uint32_t j) { return eager_mix(w + (uint64_t)(j + 1) * 0x9E3779B97F4A7C15ull); }

constexpr uint32_t kReadyRing = 4096;  // entries; entry i of the stream sits at i % kReadyRing
struct ReadyEntry {                    // one 8-byte store
  uint32_t stamp;                      // stream index + 1 (0 = never written)
  uint16_t slot;
  uint16_t events;                     // kEv* bits now set for the pair (0 = went idle)
};
struct SvcPollState {                  // device memory
  uint32_t hi_slot;                    // scan slots [0, hi_slot)
  uint32_t stop;
  uint32_t ready_next;                 // next stream index of the ready ring
  uint32_t scans;                      // completed scans (liveness)
};
struct SvcParams {
  PairDev* pairs;
  PairSvc* psvc;
  SvcCmd* cmds;        // [nowners][kOwnQ], pinned
  SvcDone* done;       // [nowners][kOwnQ], pinned
  BigBox* boxes;       // [nowners][kOwnBoxes], device
  EagerRec* erec;      // [kMaxPairs], pinned
  uint8_t* eslots;     // [kMaxPairs][kEagerMax], pinned
  SvcPollState* ps;
  uint32_t* last_ev;
  ReadyEntry* ready;
  uint32_t* host_scans;
  int nowners, nbig;
};

// launch wrappers (b200_kernels.cu)
void launch_send(PairDev* pairs, const SendOpDev* ops, OpResult* results, int nops, void* stream);
void launch_recv(PairDev* pairs, const RecvOpDev* ops, OpResult* results, int nops, void* stream);
void launch_poll_scan(PairDev* pairs, const int32_t* slots, uint32_t* events, uint32_t* ready_count,
                      int32_t* ready_slots, int n, void* stream);

// owners / pool / poller on three streams; returns false when the resident grids cannot be co-resident
bool launch_service(const SvcParams& sp, void* s_owner, void* s_big, void* s_poll);
int svc_trace_read(unsigned long long* out16);  // -DB200_SVC_TRACE builds only (device-side phase timers)

void launch_probe_copy(uint8_t* dst, const uint8_t* src, uint64_t bytes_per_cta, uint64_t stride, int nctas,
                       int threads, uint32_t mis, uint32_t item_bytes, uint32_t dynamic, void* stream);

}  // namespace b200
