// C-ABI harness around the reference's OWN classes, compiled unmodified from
// /root/reference (see oracle/Makefile): RingBufferPollable (ring_buffer.cc),
// PairPollable (pair.cc), Poller (poller.cc) over the loopback fake verbs in
// oracle/shim/.  Output: oracle/_ref/libref_pair{,_dbg}.so.
//
// TEST INFRASTRUCTURE ONLY: used by tests/ to pin oracle/rb_oracle.c and to
// generate tests/golden/, and by bench.py as the "reference" CPU baseline.
// Built with -fno-access-control so state can be inspected without touching
// the reference sources.
#include <grpc/slice.h>
#include <pthread.h>
#include <time.h>

#include <array>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "src/core/lib/ibverbs/config.h"
#include "src/core/lib/ibverbs/pair.h"
#include "src/core/lib/ibverbs/poller.h"
#include "src/core/lib/ibverbs/ring_buffer.h"

using grpc_core::ibverbs::Config;
using grpc_core::ibverbs::PairPollable;
using grpc_core::ibverbs::PairStatus;
using grpc_core::ibverbs::Poller;
using grpc_core::ibverbs::RingBufferPollable;

namespace {
struct hslice {
  const uint8_t* ptr;
  uint64_t len;
};

void to_grpc_slices(const hslice* in, size_t n, std::vector<grpc_slice>& out) {
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    memset(&out[i], 0, sizeof(grpc_slice));
    out[i].refcount = reinterpret_cast<grpc_slice_refcount*>(0x1);  // "refcounted" view: ptr+len
    out[i].data.refcounted.length = in[i].len;
    out[i].data.refcounted.bytes = const_cast<uint8_t*>(in[i].ptr);
  }
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- config
void ref_set_ring_kb(uint32_t kb) {
  auto& c = Config::Get();
  c.ring_buffer_size_kb_ = kb;
  c.zerocopy_buffer_size_kb_ = 1;  // unused by Send(); keep the footprint small
}
uint32_t ref_get_ring_kb() { return Config::Get().get_ring_buffer_size_kb(); }

// ------------------------------------------------------------------ pair
void* ref_pair_create() {
  auto* p = new PairPollable();
  p->Init();
  return p;
}
void ref_pair_destroy(void* p) { delete static_cast<PairPollable*>(p); }

// Connect() blocks until the peer's MR blob arrives, so both ends must run
// concurrently (in gRPC they are two processes).
int ref_pair_connect(void* a, void* b) {
  auto* pa = static_cast<PairPollable*>(a);
  auto* pb = static_cast<PairPollable*>(b);
  auto addr_a = pa->get_self_address().bytes();
  auto addr_b = pb->get_self_address().bytes();
  bool ok_b = false;
  std::thread t([&] { ok_b = pb->Connect(addr_a); });
  bool ok_a = pa->Connect(addr_b);
  t.join();
  return ok_a && ok_b;
}
// one side of Connect(): blocks until the peer's memory-region blob arrives, like the real thing
int ref_pair_connect_to(void* p, const void* addr, size_t n) {
  std::vector<char> v(static_cast<const char*>(addr), static_cast<const char*>(addr) + n);
  return static_cast<PairPollable*>(p)->Connect(v) ? 1 : 0;
}
const char* ref_pair_error(void* p) { return static_cast<PairPollable*>(p)->get_error().c_str(); }
size_t ref_pair_address(void* p, void* out, size_t cap) {
  auto b = static_cast<PairPollable*>(p)->get_self_address().bytes();
  if (cap >= b.size()) memcpy(out, b.data(), b.size());
  return b.size();
}
uint64_t ref_pair_send(void* p, const hslice* slices, size_t n, size_t byte_idx) {
  std::vector<grpc_slice> gs;
  to_grpc_slices(slices, n, gs);
  return static_cast<PairPollable*>(p)->Send(gs.data(), n, byte_idx);
}
uint64_t ref_pair_recv(void* p, void* dst, uint64_t cap) {
  return static_cast<PairPollable*>(p)->Recv(dst, cap);
}
int ref_pair_has_message(void* p) { return static_cast<PairPollable*>(p)->HasMessage(); }
int ref_pair_has_pending_writes(void* p) { return static_cast<PairPollable*>(p)->HasPendingWrites(); }
uint64_t ref_pair_readable(void* p) { return static_cast<PairPollable*>(p)->GetReadableSize(); }
uint64_t ref_pair_writable(void* p) { return static_cast<PairPollable*>(p)->GetWritableSize(); }
int ref_pair_get_status(void* p) { return static_cast<int>(static_cast<PairPollable*>(p)->get_status()); }
void ref_pair_disconnect(void* p) { static_cast<PairPollable*>(p)->Disconnect(); }
int ref_pair_wakeup_fd(void* p) { return static_cast<PairPollable*>(p)->get_wakeup_fd()->read_fd; }
int ref_pair_max_sge(void* p) { return static_cast<PairPollable*>(p)->max_sge_num_; }

// out[0..8) = head, moving_head, remain, remote_tail, internal_read_size,
//             partial_write, credit remote_head (as seen by this sender), peer_exit
void ref_pair_state(void* p, uint64_t* out) {
  auto* pp = static_cast<PairPollable*>(p);
  out[0] = pp->ring_buf_.head_.load();
  out[1] = pp->ring_buf_.moving_head_;
  out[2] = pp->ring_buf_.remain_.load();
  out[3] = pp->remote_tail_;
  out[4] = pp->internal_read_size_;
  out[5] = pp->partial_write_.load();
  auto* st = reinterpret_cast<PairPollable::status_report*>(
      pp->recv_buffers_[PairPollable::kStatusBuffer]->data());
  out[6] = st->remote_head;
  out[7] = static_cast<uint64_t>(st->peer_exit);
}
uint8_t* ref_pair_ring(void* p) { return static_cast<PairPollable*>(p)->ring_buf_.buf_; }
uint64_t ref_pair_ring_size(void* p) { return static_cast<PairPollable*>(p)->ring_buf_.capacity_; }
uint8_t* ref_pair_staging(void* p) {
  return static_cast<PairPollable*>(p)->send_buffers_[PairPollable::kDataBuffer]->data();
}

// rdma_flush / rdma_handle_write loop (rdma_bp_posix.cc:470-557): keep calling
// Send while it accepts bytes, advancing the slice/byte cursor as :480-493.
uint64_t ref_pair_send_all(void* p, const hslice* slices, size_t n, size_t byte_idx, uint64_t* calls) {
  std::vector<grpc_slice> gs;
  to_grpc_slices(slices, n, gs);
  auto* pp = static_cast<PairPollable*>(p);
  uint64_t total = 0, ncalls = 0;
  size_t idx = 0;
  while (idx < n) {
    uint64_t sent = pp->Send(gs.data() + idx, n - idx, byte_idx);
    if (sent == 0) break;
    ncalls++;
    total += sent;
    while (sent > 0) {
      uint64_t left = slices[idx].len - byte_idx;
      if (sent >= left) {
        sent -= left;
        idx++;
        byte_idx = 0;
      } else {
        byte_idx += sent;
        sent = 0;
      }
    }
  }
  if (calls) *calls = ncalls;
  return total;
}
// rdma_do_read loop (rdma_bp_posix.cc:180-286)
uint64_t ref_pair_recv_drain(void* p, void* dst, uint64_t cap, uint64_t* calls) {
  auto* pp = static_cast<PairPollable*>(p);
  uint64_t got = 0, ncalls = 0;
  while (got < cap) {
    uint64_t n = pp->Recv(static_cast<uint8_t*>(dst) + got, cap - got);
    if (n == 0) break;
    got += n;
    ncalls++;
  }
  if (calls) *calls = ncalls;
  return got;
}

// ---------------------------------------------------------------- poller
void ref_poller_add(void* p) { Poller::Get().AddPollable(static_cast<PairPollable*>(p)); }
void ref_poller_remove(void* p) { Poller::Get().RemovePollable(static_cast<PairPollable*>(p)); }

// ------------------------------------------------------------- ring only
// RingBufferPollable on a caller-owned buffer (any power-of-two size >= 32).
void* ref_ring_create(uint8_t* buf, uint64_t size) {
  auto* r = new RingBufferPollable(buf, size);
  r->Init();
  return r;
}
void ref_ring_destroy(void* r) { delete static_cast<RingBufferPollable*>(r); }
int ref_ring_has_message(void* r) { return static_cast<RingBufferPollable*>(r)->HasMessage(); }
uint64_t ref_ring_readable(void* r) { return static_cast<RingBufferPollable*>(r)->GetReadableSize(); }
uint64_t ref_ring_read(void* r, void* dst, uint64_t cap, uint64_t* internal) {
  return static_cast<RingBufferPollable*>(r)->Read(dst, cap, internal);
}
void ref_ring_state(void* r, uint64_t* out) {
  auto* rr = static_cast<RingBufferPollable*>(r);
  out[0] = rr->head_.load();
  out[1] = rr->moving_head_;
  out[2] = rr->remain_.load();
}
uint64_t ref_encoded_size(uint64_t p) { return RingBufferPollable::GetEncodedSize(p); }
uint64_t ref_calc_writable(uint64_t s) { return RingBufferPollable::CalculateWritableSize(s); }
uint64_t ref_free_size(void* r, uint64_t head, uint64_t tail) {
  return static_cast<RingBufferPollable*>(r)->GetFreeSize(head, tail);
}
// Writer side exactly as PairPollable::Send drives it: encode `n` frames into
// `staging` with AppendHeader/Payload/Footer, build SGEs, let GetWriteRequests
// split at the ring end, then perform the 1-2 "RDMA writes" as memcpy.
// Returns the new tail; *num_wrs = WRs used.
uint64_t ref_ring_write_frames(void* r, uint64_t tail, uint8_t* staging, const hslice* frames,
                               size_t n, int* num_wrs) {
  auto* rr = static_cast<RingBufferPollable*>(r);
  std::vector<ibv_sge> sges;
  uint64_t st = 0;
  for (size_t i = 0; i < n; i++) {
    uint8_t* base = staging + st;
    uint8_t* q = RingBufferPollable::AppendHeader(base, frames[i].len);
    q = RingBufferPollable::AppendPayload(q, const_cast<uint8_t*>(frames[i].ptr), frames[i].len);
    q = RingBufferPollable::AppendFooter(q);
    ibv_sge s;
    s.addr = reinterpret_cast<uint64_t>(base);
    s.length = static_cast<uint32_t>(q - base);
    s.lkey = 0;
    sges.push_back(s);
    st += s.length;
  }
  std::array<ibv_send_wr, 2> wrs;
  uint64_t new_tail = rr->GetWriteRequests(tail, rr->buf_, 0, sges, wrs);
  int used = 0;
  for (ibv_send_wr* w = &wrs[0]; w != nullptr; w = w->next) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(w->wr.rdma.remote_addr);
    for (int i = 0; i < w->num_sge; i++) {
      memcpy(dst, reinterpret_cast<void*>(w->sg_list[i].addr), w->sg_list[i].length);
      dst += w->sg_list[i].length;
    }
    used++;
  }
  if (num_wrs) *num_wrs = used;
  return new_tail;
}

// ------------------------------------------------- multi-threaded baseline
// Same workload shape as orb_bench_stream (oracle/rb_oracle.c) but through the
// reference's PairPollable::Send/Recv.
struct ref_worker {
  int first_conn, n_conn, warm, msgs;
  const uint64_t* lens;
  size_t nslices;
  uint64_t delivered, checksum;
  pthread_barrier_t* bar;
};

static uint64_t fnv1a(const uint8_t* p, uint64_t n, uint64_t h) {
  for (uint64_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}

static void* ref_worker_main(void* arg) {
  auto* w = static_cast<ref_worker*>(arg);
  uint64_t msg_bytes = 0;
  for (size_t i = 0; i < w->nslices; i++) msg_bytes += w->lens[i];
  int nc = w->n_conn;
  std::vector<PairPollable*> tx(nc), rx(nc);
  std::vector<std::vector<uint8_t>> src(nc), dst(nc);
  std::vector<hslice> sl(w->nslices);
  std::vector<grpc_slice> gs;
  for (int c = 0; c < nc; c++) {
    tx[c] = static_cast<PairPollable*>(ref_pair_create());
    rx[c] = static_cast<PairPollable*>(ref_pair_create());
    ref_pair_connect(tx[c], rx[c]);
    src[c].resize(msg_bytes);
    dst[c].assign(msg_bytes, 0);
    for (uint64_t i = 0; i < msg_bytes; i++)
      src[c][i] = static_cast<uint8_t>(i + 131u * static_cast<unsigned>(w->first_conn + c));
  }
  uint64_t delivered = 0;
  for (int m = -w->warm; m < w->msgs; m++) {
    if (m == 0) {
      pthread_barrier_wait(w->bar);  // warm-up messages done: timing starts
      delivered = 0;
    }
    for (int c = 0; c < nc; c++) {
      uint64_t off = 0;
      for (size_t i = 0; i < w->nslices; i++) {
        sl[i].ptr = src[c].data() + off;
        sl[i].len = w->lens[i];
        off += w->lens[i];
      }
      to_grpc_slices(sl.data(), w->nslices, gs);
      size_t idx = 0, bidx = 0;
      uint64_t got = 0;
      while (got < msg_bytes) {
        while (idx < w->nslices) {
          uint64_t sent = tx[c]->Send(gs.data() + idx, w->nslices - idx, bidx);
          if (sent == 0) break;
          while (sent > 0) {
            uint64_t left = sl[idx].len - bidx;
            if (sent >= left) { sent -= left; idx++; bidx = 0; }
            else { bidx += sent; sent = 0; }
          }
        }
        uint64_t n;
        while (got < msg_bytes && (n = rx[c]->Recv(dst[c].data() + got, msg_bytes - got)) > 0) got += n;
      }
      delivered += got;
    }
  }
  pthread_barrier_wait(w->bar);
  uint64_t h = 0xcbf29ce484222325ULL;
  for (int c = 0; c < nc; c++) h = fnv1a(dst[c].data(), msg_bytes, h);
  w->delivered = delivered;
  w->checksum = h;
  for (int c = 0; c < nc; c++) {
    tx[c]->Disconnect();
    rx[c]->Disconnect();
    delete tx[c];
    delete rx[c];
  }
  return nullptr;
}

// ------------------------------------------------------------ unary ping-pong
// BASELINE config 3 at the pair level: `conns` connections served by `groups` client threads and
// `groups` server threads (client i and server i own the same conns/groups connections).  A client
// sends `msg_bytes` on a connection, spins on HasMessage, receives the echo, then moves to its
// next connection; a server polls its connections round-robin (the Poller / busy-poll loop),
// receives and echoes.  rtt_ns[c * iters + k] = round-trip time of iteration k on connection c.
struct pp_worker {
  int first_conn, n_conn, iters, warm, is_server;
  uint64_t msg_bytes;
  PairPollable** cli;
  PairPollable** srv;
  uint64_t* rtt_ns;
  pthread_barrier_t* bar;
  std::atomic<int>* stop;
};

static inline uint64_t now_ns() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return static_cast<uint64_t>(t.tv_sec) * 1000000000ull + static_cast<uint64_t>(t.tv_nsec);
}

static void* pp_main(void* arg) {
  auto* w = static_cast<pp_worker*>(arg);
  std::vector<uint8_t> buf(w->msg_bytes), in(w->msg_bytes + 64);
  for (uint64_t i = 0; i < w->msg_bytes; i++) buf[i] = static_cast<uint8_t>(i * 7 + w->first_conn);
  hslice hs{buf.data(), w->msg_bytes};
  std::vector<grpc_slice> gs;
  to_grpc_slices(&hs, 1, gs);
  pthread_barrier_wait(w->bar);
  if (w->is_server) {
    std::vector<uint64_t> got(w->n_conn, 0);
    while (!w->stop->load(std::memory_order_relaxed)) {
      for (int c = 0; c < w->n_conn; c++) {
        PairPollable* p = w->srv[w->first_conn + c];
        if (!p->HasMessage()) continue;
        uint64_t n = p->Recv(in.data(), w->msg_bytes - got[c]);
        got[c] += n;
        if (got[c] == w->msg_bytes) {
          got[c] = 0;
          hslice es{in.data(), w->msg_bytes};
          std::vector<grpc_slice> eg;
          to_grpc_slices(&es, 1, eg);
          uint64_t sent = 0;
          while (sent < w->msg_bytes) sent += p->Send(eg.data(), 1, sent);
        }
      }
    }
  } else {
    for (int k = -w->warm; k < w->iters; k++) {
      for (int c = 0; c < w->n_conn; c++) {
        PairPollable* p = w->cli[w->first_conn + c];
        const uint64_t t0 = now_ns();
        uint64_t sent = 0;
        while (sent < w->msg_bytes) sent += p->Send(gs.data(), 1, sent);
        uint64_t got = 0;
        while (got < w->msg_bytes) {
          while (!p->HasMessage()) {
          }
          got += p->Recv(in.data(), w->msg_bytes - got);
        }
        if (k >= 0) w->rtt_ns[static_cast<size_t>(w->first_conn + c) * w->iters + k] = now_ns() - t0;
      }
    }
  }
  return nullptr;
}

double ref_bench_pingpong(int conns, int groups, int iters, int warm, uint64_t msg_bytes, uint64_t ring_capacity,
                          uint64_t* rtt_ns) {
  ref_set_ring_kb(static_cast<uint32_t>(ring_capacity / 1024));
  if (groups < 1) groups = 1;
  if (groups > conns) groups = conns;
  std::vector<PairPollable*> cli(conns), srv(conns);
  for (int c = 0; c < conns; c++) {
    cli[c] = static_cast<PairPollable*>(ref_pair_create());
    srv[c] = static_cast<PairPollable*>(ref_pair_create());
    ref_pair_connect(cli[c], srv[c]);
  }
  std::atomic<int> stop{0};
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, 2 * groups + 1);
  std::vector<pthread_t> th(2 * groups);
  std::vector<pp_worker> ws(2 * groups);
  int base = 0;
  for (int g = 0; g < groups; g++) {
    int nc = conns / groups + (g < conns % groups ? 1 : 0);
    ws[2 * g] = pp_worker{base, nc, iters, warm, 0, msg_bytes, cli.data(), srv.data(), rtt_ns, &bar, &stop};
    ws[2 * g + 1] = pp_worker{base, nc, iters, warm, 1, msg_bytes, cli.data(), srv.data(), rtt_ns, &bar, &stop};
    base += nc;
    pthread_create(&th[2 * g], nullptr, pp_main, &ws[2 * g]);
    pthread_create(&th[2 * g + 1], nullptr, pp_main, &ws[2 * g + 1]);
  }
  pthread_barrier_wait(&bar);
  const uint64_t t0 = now_ns();
  for (int g = 0; g < groups; g++) pthread_join(th[2 * g], nullptr);  // clients finish
  const uint64_t t1 = now_ns();
  stop.store(1);
  for (int g = 0; g < groups; g++) pthread_join(th[2 * g + 1], nullptr);
  pthread_barrier_destroy(&bar);
  for (int c = 0; c < conns; c++) {
    cli[c]->Disconnect();
    srv[c]->Disconnect();
    delete cli[c];
    delete srv[c];
  }
  return 1e-9 * static_cast<double>(t1 - t0);
}

double ref_bench_stream(int conns, int threads, int warm, int msgs, uint64_t ring_capacity,
                        const uint64_t* lens, size_t nslices, uint64_t* delivered,
                        uint64_t* checksum) {
  ref_set_ring_kb(static_cast<uint32_t>(ring_capacity / 1024));
  if (threads < 1) threads = 1;
  if (threads > conns) threads = conns;
  std::vector<pthread_t> th(threads);
  std::vector<ref_worker> ws(threads);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, threads + 1);
  int base = 0;
  for (int t = 0; t < threads; t++) {
    int nc = conns / threads + (t < conns % threads ? 1 : 0);
    ws[t] = ref_worker{base, nc, warm, msgs, lens, nslices, 0, 0, &bar};
    base += nc;
    pthread_create(&th[t], nullptr, ref_worker_main, &ws[t]);
  }
  timespec t0, t1;
  pthread_barrier_wait(&bar);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&bar);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  uint64_t d = 0, h = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], nullptr);
    d += ws[t].delivered;
    h ^= ws[t].checksum;
  }
  pthread_barrier_destroy(&bar);
  if (delivered) *delivered = d;
  if (checksum) *checksum = h;
  return static_cast<double>(t1.tv_sec - t0.tv_sec) + 1e-9 * static_cast<double>(t1.tv_nsec - t0.tv_nsec);
}

}  // extern "C"
