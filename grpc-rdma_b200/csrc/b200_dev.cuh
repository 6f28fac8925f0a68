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

// ---- persistent service kernel (k_service): host -> device commands in pinned mapped memory.
// One command slot per worker CTA; the host writes the fields, then bumps `seq` (release); the
// CTA polls `seq` (acquire, system scope), runs the op with the same code as k_send / k_recv and
// answers in SvcDone.  The last CTA of the grid is the poller: it scans the connection table
// continuously, keeps the host-visible mirrors current and appends readiness CHANGES to the
// ready ring (warp-aggregated: one atomic per warp).
constexpr uint32_t kSvcSend = 1, kSvcRecv = 2, kSvcStop = 3;
constexpr uint32_t kSvcInline = 5;  // slices carried inside the command (a unary call has 2-4)
struct __align__(128) SvcCmd {
  uint32_t seq;      // command number, written last by the host
  uint32_t op;       // kSvcSend / kSvcRecv / kSvcStop
  int32_t slot;
  uint32_t flags;    // B200_BATCH_*
  uint64_t ptr;      // send: SliceDev* (GPU-addressable; unused when n <= kSvcInline)   recv: destination
  uint64_t n;        // send: nslices                        recv: capacity
  uint64_t byte_idx;
  SliceDev inl[kSvcInline];  // send: the slice list itself when it is short -- no second trip over PCIe
};
static_assert(sizeof(SvcCmd) == 128, "one command = one 128-byte line");
struct __align__(32) SvcDone {
  uint64_t bytes, calls;
  uint32_t seq;      // = SvcCmd.seq once the op is finished and its bytes are visible
  uint32_t _pad[3];
};
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

// launch wrappers (b200_kernels.cu)
void launch_send(PairDev* pairs, const SendOpDev* ops, OpResult* results, int nops, void* stream);
void launch_recv(PairDev* pairs, const RecvOpDev* ops, OpResult* results, int nops, void* stream);
void launch_poll_scan(PairDev* pairs, const int32_t* slots, uint32_t* events, uint32_t* ready_count,
                      int32_t* ready_slots, int n, void* stream);

void launch_service(PairDev* pairs, SvcCmd* cmds, SvcDone* done, SvcPollState* ps, uint32_t* last_ev,
                    ReadyEntry* ready, uint32_t* host_scans, int nworkers, void* stream);

void launch_probe_copy(uint8_t* dst, const uint8_t* src, uint64_t bytes_per_cta, uint64_t stride, int nctas,
                       int threads, uint32_t mis, uint32_t item_bytes, uint32_t dynamic, void* stream);

}  // namespace b200
