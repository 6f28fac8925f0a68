// b200_endpoint.cc -- endpoint state machine + BPEV hybrid poll loop over the pair ABI.
// See include/b200_endpoint.h for the reference functions each piece mirrors
// (src/core/lib/iomgr/rdma_bp_posix.cc, ev_epollex_rdma_bpev_linux.cc).
//
// Two ways of running the pair operations:
//   * call by call (ops->send / ops->recv from the closures, exactly the reference's structure) -- used when
//     the ops table has no `submit` (the CPU tables of the tests: oracle, the reference's own PairPollable);
//   * batched (ops->submit present: the CUDA library): rdma_read / rdma_write / the readiness scan only
//     QUEUE the endpoint, and one b200_engine_work pass executes every queued rdma_flush loop and
//     rdma_do_read loop with ONE submit -- all connections' ops run on the GPU side by side instead of one
//     round trip after the other.  Read slices come from a pool of pinned, 256-byte aligned blocks and
//     their size adapts to what the last reads delivered (what gRPC's own tcp_posix.cc does with
//     target_length; the reference's max(256, GetReadableSize()) is the starting point and the floor).
#include "../../include/b200_endpoint.h"

#include <errno.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

constexpr int kMaxEpollEvents = 100;  // MAX_EPOLL_EVENTS, ev_epollex_rdma_bpev_linux.cc
constexpr size_t kMaxReadIovec = 4;   // MAX_READ_IOVEC, rdma_bp_posix.cc:178
constexpr intptr_t kTagPair = 2;      // eventfd of a pair, tag ptr|2 (:725-741)
constexpr intptr_t kTagWake = 4;      // the engine's own wakeup fd

// ---- default pair ops: the CUDA library
void* d_take(const char* id) { return b200_pool_take(id); }
void d_putback(void* p) { b200_pool_putback((b200_pair*)p); }
void d_init(void* p) { b200_pair_init((b200_pair*)p); }
size_t d_addr(void* p, void* o) { return b200_pair_self_address((b200_pair*)p, o); }
int d_connect(void* p, const void* a, size_t n) { return b200_pair_connect((b200_pair*)p, a, n); }
uint64_t d_send(void* p, const b200_slice* s, size_t n, size_t b) { return b200_pair_send((b200_pair*)p, s, n, b); }
uint64_t d_recv(void* p, void* d, uint64_t c) { return b200_pair_recv((b200_pair*)p, d, c); }
int d_has_msg(const void* p) { return b200_pair_has_message((const b200_pair*)p); }
int d_pending(const void* p) { return b200_pair_has_pending_writes((const b200_pair*)p); }
uint64_t d_readable(const void* p) { return b200_pair_readable((const b200_pair*)p); }
int d_status(void* p) { return (int)b200_pair_status((b200_pair*)p); }
const char* d_error(const void* p) { return b200_pair_error((const b200_pair*)p); }
int d_wfd(void* p) { return b200_pair_wakeup_read_fd((b200_pair*)p); }
void d_consume(void* p) { b200_pair_consume_wakeup((b200_pair*)p); }
void d_disconnect(void* p) { b200_pair_disconnect((b200_pair*)p); }
void d_padd(void* p) { b200_poller_add((b200_pair*)p); }
void d_premove(void* p) { b200_poller_remove((b200_pair*)p); }
// batching needs the resident service kernels (no launch per pass, slices used in place); without them the
// engine runs the queued continuations call by call (rc 1 = "not available")
int d_submit(const b200_send_op* s, size_t ns, uint64_t* acc, const b200_recv_op* r, size_t nr, uint64_t* del, int flags) {
  if (!b200_service_running()) return 1;
  return b200_pairs_submit(s, ns, acc, r, nr, del, flags);
}
void* d_post_send(void* p, const b200_slice* s, size_t n, size_t b, int flags, int* again) {
  if (!b200_service_running()) {
    *again = 2;
    return nullptr;
  }
  return b200_pair_post_send((b200_pair*)p, s, n, b, flags, again);
}
void* d_post_recv(void* p, void* dst, uint64_t cap, int flags, int* again) {
  if (!b200_service_running()) {
    *again = 2;
    return nullptr;
  }
  return b200_pair_post_recv((b200_pair*)p, dst, cap, flags, again);
}
int d_poll(void* op, uint64_t* bytes) { return b200_async_poll((b200_async*)op, bytes); }
void* d_malloc(size_t n) { return b200_mem_alloc_host(n); }
void d_mfree(void* p) { b200_mem_free_host(p); }
const b200_pair_ops kCudaOps = {d_take,   d_putback, d_init,       d_addr,     d_connect, d_send,
                                d_recv,   d_has_msg, d_pending,    d_readable, d_status,  d_error,
                                d_wfd,    d_consume, d_disconnect, d_padd,     d_premove, d_submit,
                                d_malloc, d_mfree,   d_post_send,  d_post_recv, d_poll};

// Pool of read-slice blocks: power-of-two size classes from 256 bytes, GPU-addressable when the ops table
// provides an allocator (every b200_mem_alloc_host block is 256-byte aligned: the alignment the PCIe DMA
// wants, profiles/r1k_pcie_alignment.txt), plain aligned host memory otherwise.
struct Pool {
  const b200_pair_ops* ops;
  std::mutex mu;
  std::vector<void*> free_[32];
  std::vector<void*> all;
  explicit Pool(const b200_pair_ops* o) : ops(o) {}
  ~Pool() {
    for (void* p : all) {
      if (ops->mem_free) ops->mem_free(p);
      else free(p);
    }
  }
  static int cls(size_t bytes) {
    int c = 8;  // 256
    while (((size_t)1 << c) < bytes) c++;
    return c;
  }
  void* get(size_t bytes, int* c_out) {
    const int c = cls(bytes);
    *c_out = c;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!free_[c].empty()) {
        void* p = free_[c].back();
        free_[c].pop_back();
        return p;
      }
    }
    void* p = nullptr;
    if (ops->mem_alloc) p = ops->mem_alloc((size_t)1 << c);
    else if (posix_memalign(&p, 256, (size_t)1 << c) != 0) p = nullptr;
    if (p) {
      std::lock_guard<std::mutex> lk(mu);
      all.push_back(p);
    }
    return p;
  }
  void put(void* p, int c) {
    std::lock_guard<std::mutex> lk(mu);
    free_[c].push_back(p);
  }
};

// A read slice: the endpoint allocates it itself (rdma_bp_posix.cc:308-317) and hands views out.
struct Buf {
  std::shared_ptr<uint8_t> store;
  size_t off = 0, len = 0;
  uint8_t* ptr() const { return store.get() + off; }
};

Buf make_buf(const std::shared_ptr<Pool>& pool, size_t bytes) {
  int c = 0;
  void* p = pool->get(bytes, &c);
  Buf b;
  if (!p) abort();
  std::shared_ptr<Pool> keep = pool;
  b.store = std::shared_ptr<uint8_t>((uint8_t*)p, [keep, c](uint8_t* q) { keep->put(q, c); });
  b.off = 0;
  b.len = bytes;
  return b;
}

// lockfree_event stand-in (grpc_fd read_closure / write_closure): an edge is remembered until
// somebody asks for it.
struct Event {
  bool ready = false;
  bool armed = false;
};

struct Task {
  std::function<void()> f;
  bool user;  // a user callback: runs WITHOUT the engine lock
};

}  // namespace

struct b200_endpoint {
  b200_engine* engine = nullptr;
  void* pair = nullptr;
  int fd = -1;
  bool enable_poller = false;
  bool is_first_read = true;
  int inq = 1;
  bool shutdown = false;
  std::string shutdown_why;
  std::string peer_string;
  int refs = 1;
  // read side
  b200_closure_fn read_cb = nullptr;
  void* read_arg = nullptr;
  std::vector<Buf> incoming;          // incoming_buffer
  std::vector<Buf> last_read_buffer;  // garbage after the last read
  std::vector<b200_slice> incoming_view;
  Event rd;
  size_t target_length = 0;  // batching: adaptive read size (tcp_posix.cc target_length)
  bool queued_read = false, queued_write = false;
  // write side
  b200_closure_fn write_cb = nullptr;
  void* write_arg = nullptr;
  const b200_slice* outgoing = nullptr;  // outgoing_buffer (caller-owned)
  size_t outgoing_count = 0;
  size_t outgoing_idx = 0;       // slices already removed from the front
  size_t outgoing_byte_idx = 0;  // byte within outgoing[outgoing_idx] to write next
  Event wr;
};

struct b200_engine {
  const b200_pair_ops* ops = nullptr;
  int epfd = -1;
  int wake_fd = -1;  // kicked when work is queued for an engine that may be asleep in epoll_wait
  int busy_poll_us = 500;
  std::mutex mu;  // rdma_mu + the pollable's own lock: engine and endpoint state; never held across
                  // epoll_wait, a submit or a user callback
  std::vector<b200_endpoint*> rdma_fds;
  std::deque<Task> exec;  // ExecCtx closure list
  bool flushing = false;
  bool batch = false;
  bool in_work = false;  // a thread is between the end of its wait and the end of its pass
  size_t max_read_chunk = 4u << 20;
  std::vector<b200_endpoint*> q_reads, q_writes;  // batching: rdma_do_read / rdma_flush loops to run
  struct InFlight {
    b200_endpoint* ep;
    bool is_read;
    void* op;
    size_t rlen;
  };
  std::vector<InFlight> inflight;  // completion-queue form: posted, not finished yet
  bool async = false;
  std::shared_ptr<Pool> pool;
  uint64_t stats[4] = {0, 0, 0, 0};
  uint64_t bstats[3] = {0, 0, 0};
};

namespace {

using Lock = std::unique_lock<std::mutex>;

void schedule(b200_engine* e, std::function<void()> f) { e->exec.push_back(Task{std::move(f), false}); }
void schedule_user(b200_engine* e, b200_closure_fn cb, void* arg, const char* error) {
  if (error) {
    std::string es(error);
    e->exec.push_back(Task{[cb, arg, es] { cb(arg, es.c_str()); }, true});
  } else {
    e->exec.push_back(Task{[cb, arg] { cb(arg, nullptr); }, true});
  }
}

// grpc_core::ExecCtx::Flush.  Called with the lock held; user callbacks run unlocked.
void flush(b200_engine* e, Lock& lk) {
  if (e->flushing) return;  // the thread that is flushing picks the new tasks up
  e->flushing = true;
  while (!e->exec.empty()) {
    Task t = std::move(e->exec.front());
    e->exec.pop_front();
    if (t.user) {
      lk.unlock();
      t.f();
      lk.lock();
    } else {
      t.f();
    }
  }
  e->flushing = false;
}

void kick_engine(b200_engine* e) {
  if (e->wake_fd >= 0) (void)eventfd_write(e->wake_fd, 1);
}

size_t buf_length(const std::vector<Buf>& v) {
  size_t n = 0;
  for (auto& b : v) n += b.len;
  return n;
}

// rdma_annotate_error, rdma_bp_posix.cc:86-96
std::string annotate(b200_endpoint* ep, const std::string& msg) {
  return msg + " [fd " + std::to_string(ep->fd) + ", grpc_status UNAVAILABLE, target_address " + ep->peer_string + "]";
}

void ep_unref(b200_endpoint* ep);
void rdma_handle_read(b200_endpoint* ep, const char* error);
void rdma_handle_write(b200_endpoint* ep, const char* error);

// the readiness edge reaches the endpoint: run (or, batching, queue) its read / write continuation
void dispatch(b200_endpoint* ep, bool is_read) {
  b200_engine* e = ep->engine;
  if (e->batch) {
    bool& q = is_read ? ep->queued_read : ep->queued_write;
    if (!q) {
      q = true;
      const bool was_idle = e->q_reads.empty() && e->q_writes.empty();
      (is_read ? e->q_reads : e->q_writes).push_back(ep);
      if (was_idle && !e->in_work) kick_engine(e);  // the engine may be asleep in epoll_wait
    }
  } else {
    schedule(e, [ep, is_read] { is_read ? rdma_handle_read(ep, nullptr) : rdma_handle_write(ep, nullptr); });
  }
}

// grpc_fd_notify_on_read / _write over the lockfree event
void notify_on(b200_endpoint* ep, Event& ev, bool is_read) {
  b200_engine* e = ep->engine;
  if (ep->shutdown) {
    std::string why = ep->shutdown_why;
    schedule(e, [ep, is_read, why] { is_read ? rdma_handle_read(ep, why.c_str()) : rdma_handle_write(ep, why.c_str()); });
    return;
  }
  if (ev.ready) {
    ev.ready = false;
    dispatch(ep, is_read);
  } else {
    ev.armed = true;
  }
}

// fd_become_readable / fd_become_writable
void set_ready(b200_endpoint* ep, Event& ev, bool is_read) {
  if (ev.armed) {
    ev.armed = false;
    dispatch(ep, is_read);
  } else {
    ev.ready = true;
  }
}

// call_read_cb, rdma_bp_posix.cc:170-176
void call_read_cb(b200_endpoint* ep, const char* error) {
  b200_closure_fn cb = ep->read_cb;
  void* arg = ep->read_arg;
  ep->read_cb = nullptr;
  ep->incoming_view.clear();
  if (!error)
    for (auto& b : ep->incoming) ep->incoming_view.push_back({b.ptr(), b.len});
  schedule_user(ep->engine, cb, arg, error);
}

// grpc_slice_buffer_trim_end(incoming, n, &last_read_buffer)
void trim_end(std::vector<Buf>& buf, size_t n, std::vector<Buf>& garbage) {
  while (n > 0 && !buf.empty()) {
    Buf& last = buf.back();
    if (last.len > n) {
      Buf tail = last;
      tail.off = last.off + (last.len - n);
      tail.len = n;
      last.len -= n;
      garbage.insert(garbage.begin(), tail);
      n = 0;
    } else {
      n -= last.len;
      garbage.insert(garbage.begin(), last);
      buf.pop_back();
    }
  }
}

// the part of rdma_do_read after its Recv loop (rdma_bp_posix.cc:216-286): `total_read_bytes` came back
void finish_read(b200_endpoint* ep, size_t total_read_bytes, size_t incoming_length) {
  const b200_pair_ops* ops = ep->engine->ops;
  if (total_read_bytes == 0) {
    ep->inq = ops->readable(ep->pair) > 0;
    const int status = ops->status(ep->pair);
    if (status == B200_HALF_CLOSED) {  // :218-228
      ep->incoming.clear();
      call_read_cb(ep, annotate(ep, "Pair closed").c_str());
      ep_unref(ep);
      return;
    } else if (status == B200_ERROR) {  // :229-239
      ep->incoming.clear();
      call_read_cb(ep, annotate(ep, std::string("Pair error, ") + ops->error(ep->pair)).c_str());
      ep_unref(ep);
      return;
    }
    notify_on(ep, ep->rd, true);  // we've consumed the edge, request a new one
    return;
  }
  if (total_read_bytes < incoming_length) trim_end(ep->incoming, incoming_length - total_read_bytes, ep->last_read_buffer);
  call_read_cb(ep, nullptr);
  ep_unref(ep);
}

// rdma_do_read, rdma_bp_posix.cc:180-286 (call-by-call form)
void rdma_do_read(b200_endpoint* ep) {
  const b200_pair_ops* ops = ep->engine->ops;
  struct Iov {
    uint8_t* base;
    size_t len;
  } iov[kMaxReadIovec];
  size_t total_read_bytes = 0;
  size_t iov_len = std::min(kMaxReadIovec, ep->incoming.size());
  const size_t incoming_length = buf_length(ep->incoming);
  for (size_t i = 0; i < iov_len; i++) iov[i] = {ep->incoming[i].ptr(), ep->incoming[i].len};
  while (true) {
    ep->inq = 1;  // assume there is something on the queue
    uint64_t read_bytes = 0;
    if (iov_len > 0) read_bytes = ops->recv(ep->pair, iov[0].base, iov[0].len);
    if (read_bytes == 0) {
      ep->inq = ops->readable(ep->pair) > 0;
      break;  // deliver what previous Recv calls got, or handle "nothing at all"
    }
    total_read_bytes += read_bytes;
    if (ep->inq == 0 || total_read_bytes == incoming_length) break;
    // partial read with space left: adjust the iovs and try to read more (:256-273)
    size_t remaining = read_bytes, j = 0;
    for (size_t i = 0; i < iov_len; i++) {
      if (remaining >= iov[i].len) {
        remaining -= iov[i].len;
        continue;
      }
      if (remaining > 0) {
        iov[j] = {iov[i].base + remaining, iov[i].len - remaining};
        remaining = 0;
      } else {
        iov[j] = iov[i];
      }
      ++j;
    }
    iov_len = j;
  }
  finish_read(ep, total_read_bytes, incoming_length);
}

// rdma_continue_read, rdma_bp_posix.cc:306-326: ONE slice of max(256, readable)
void prepare_read_slice(b200_endpoint* ep) {
  b200_engine* e = ep->engine;
  size_t target = std::max<uint64_t>(256, e->ops->readable(ep->pair));
  if (e->batch) target = std::min(e->max_read_chunk, std::max(target, ep->target_length));
  if (buf_length(ep->incoming) == 0 && ep->incoming.size() < kMaxReadIovec) ep->incoming.push_back(make_buf(e->pool, target));
}
void rdma_continue_read(b200_endpoint* ep) {
  prepare_read_slice(ep);
  rdma_do_read(ep);
}

// rdma_handle_read, rdma_bp_posix.cc:328-341
void rdma_handle_read(b200_endpoint* ep, const char* error) {
  if (error) {
    ep->incoming.clear();
    ep->last_read_buffer.clear();
    call_read_cb(ep, error);
    ep_unref(ep);
  } else {
    rdma_continue_read(ep);
  }
}

// the part of rdma_flush after Send (rdma_bp_posix.cc:480-524): `sent` bytes were accepted.
// true = finished (ok or error), false = partial.
bool finish_flush(b200_endpoint* ep, uint64_t sent, std::string* error) {
  const b200_pair_ops* ops = ep->engine->ops;
  size_t idx = ep->outgoing_idx;
  while (sent > 0) {  // :480-493
    const uint64_t slice_len = ep->outgoing[idx].len - ep->outgoing_byte_idx;
    if (sent >= slice_len) {
      sent -= slice_len;
      idx++;
      ep->outgoing_byte_idx = 0;
    } else {
      ep->outgoing_byte_idx += sent;
      break;
    }
  }
  if (ep->outgoing_byte_idx > 0 || idx < ep->outgoing_count) {  // partial send
    const int status = ops->status(ep->pair);
    if (status == B200_CONNECTED) {
      ep->outgoing_idx = idx;  // grpc_slice_buffer_remove_first for everything fully written
      return false;
    }
    if (status == B200_HALF_CLOSED) *error = annotate(ep, "Peer has been exited");
    else *error = annotate(ep, std::string("RDMA Pair has an internal error, ") + ops->error(ep->pair));
    return true;
  }
  ep->outgoing_idx = idx;
  error->clear();
  return true;
}

// rdma_flush, rdma_bp_posix.cc:470-524
bool rdma_flush(b200_endpoint* ep, std::string* error) {
  const size_t idx = ep->outgoing_idx;
  const uint64_t sent = ep->engine->ops->send(ep->pair, ep->outgoing + idx, ep->outgoing_count - idx, ep->outgoing_byte_idx);
  return finish_flush(ep, sent, error);
}

void finish_write(b200_endpoint* ep, const char* err) {
  b200_closure_fn cb = ep->write_cb;
  void* arg = ep->write_arg;
  ep->write_cb = nullptr;
  schedule_user(ep->engine, cb, arg, err);
  ep_unref(ep);
}

// rdma_handle_write, rdma_bp_posix.cc:527-557
void rdma_handle_write(b200_endpoint* ep, const char* error) {
  if (error) {
    finish_write(ep, error);
    return;
  }
  std::string err;
  if (!rdma_flush(ep, &err)) notify_on(ep, ep->wr, false);  // "write: delayed"
  else finish_write(ep, err.empty() ? nullptr : err.c_str());
}

// rdma_free, rdma_bp_posix.cc:112-132
void ep_free(b200_endpoint* ep) {
  b200_engine* e = ep->engine;
  const b200_pair_ops* ops = e->ops;
  e->q_reads.erase(std::remove(e->q_reads.begin(), e->q_reads.end(), ep), e->q_reads.end());
  e->q_writes.erase(std::remove(e->q_writes.begin(), e->q_writes.end(), ep), e->q_writes.end());
  if (ep->pair) {
    // fd_orphan: drop the eventfd from the epoll set and the fd from the rdma list
    epoll_ctl(e->epfd, EPOLL_CTL_DEL, ops->wakeup_read_fd(ep->pair), nullptr);
    e->rdma_fds.erase(std::remove(e->rdma_fds.begin(), e->rdma_fds.end(), ep), e->rdma_fds.end());
    if (ep->enable_poller) ops->poller_remove(ep->pair);
    ops->disconnect(ep->pair);
    ops->pool_putback(ep->pair);
    ep->pair = nullptr;
  }
  if (ep->fd >= 0) close(ep->fd);
  delete ep;
}

void ep_unref(b200_endpoint* ep) {
  if (--ep->refs == 0) ep_free(ep);
}

void complete_write(b200_endpoint* ep, uint64_t accepted);
void complete_read(b200_endpoint* ep, size_t got, size_t rlen);

// Batching: every queued rdma_flush loop and rdma_do_read loop of this pass in ONE submit.  The lock is
// released while the GPU works; the queued endpoints are kept alive by the reference their pending
// read / write holds.
void run_batch(b200_engine* e, Lock& lk) {
  if (e->q_reads.empty() && e->q_writes.empty()) return;
  std::vector<b200_endpoint*> reads, writes;
  reads.swap(e->q_reads);
  writes.swap(e->q_writes);
  std::vector<b200_send_op> sops;
  std::vector<b200_recv_op> rops;
  std::vector<b200_endpoint*> seps, reps;
  std::vector<size_t> rlen;
  for (b200_endpoint* ep : writes) {
    ep->queued_write = false;
    if (!ep->write_cb) continue;
    b200_send_op o;
    o.pair = (b200_pair*)ep->pair;
    o.slices = ep->outgoing + ep->outgoing_idx;
    o.nslices = ep->outgoing_count - ep->outgoing_idx;
    o.byte_idx = ep->outgoing_byte_idx;
    sops.push_back(o);
    seps.push_back(ep);
  }
  for (b200_endpoint* ep : reads) {
    ep->queued_read = false;
    if (!ep->read_cb) continue;
    prepare_read_slice(ep);
    ep->inq = 1;
    b200_recv_op o;
    o.pair = (b200_pair*)ep->pair;
    o.dst = ep->incoming.empty() ? nullptr : ep->incoming[0].ptr();
    o.cap = ep->incoming.empty() ? 0 : ep->incoming[0].len;
    rops.push_back(o);
    reps.push_back(ep);
    rlen.push_back(buf_length(ep->incoming));
  }
  if (sops.empty() && rops.empty()) return;
  std::vector<uint64_t> acc(sops.size() + 1, 0), del(rops.size() + 1, 0);
  e->bstats[0]++;
  e->bstats[1] += sops.size();
  e->bstats[2] += rops.size();
  lk.unlock();
  const int src = e->ops->submit(sops.data(), sops.size(), acc.data(), rops.data(), rops.size(), del.data(),
                                 B200_BATCH_UNTIL_BLOCKED);
  lk.lock();
  if (src == 1) {  // no batching right now: the same continuations, call by call
    for (b200_endpoint* ep : seps) rdma_handle_write(ep, nullptr);
    for (b200_endpoint* ep : reps) rdma_do_read(ep);
    return;
  }
  for (size_t i = 0; i < seps.size(); i++) complete_write(seps[i], acc[i]);
  for (size_t i = 0; i < reps.size(); i++) complete_read(reps[i], (size_t)del[i], rlen[i]);
}

void complete_write(b200_endpoint* ep, uint64_t accepted) {
  std::string err;
  if (!finish_flush(ep, accepted, &err)) notify_on(ep, ep->wr, false);
  else finish_write(ep, err.empty() ? nullptr : err.c_str());
}
void complete_read(b200_endpoint* ep, size_t got, size_t rlen) {
  b200_engine* e = ep->engine;
  if (got) {
    ep->inq = e->ops->readable(ep->pair) > 0;
    // adapt the next read to what this one found (tcp_posix.cc: finish_estimate / target_length)
    const size_t cap = ep->incoming.empty() ? 0 : ep->incoming[0].len;
    if (got == cap) ep->target_length = std::min(e->max_read_chunk, std::max<size_t>(2 * cap, 4096));
    else if (got < cap / 2) ep->target_length = std::max<size_t>(256, cap / 2);
  }
  finish_read(ep, got, rlen);
}

// Completion-queue form of run_batch: poll what is in flight, post what is queued; never waits for the GPU.
// Returns false when posting is not available right now (no service): the caller runs the batch path.
bool pump(b200_engine* e) {
  const b200_pair_ops* ops = e->ops;
  for (size_t i = 0; i < e->inflight.size();) {
    b200_engine::InFlight f = e->inflight[i];
    uint64_t bytes = 0;
    const int r = ops->poll(f.op, &bytes);
    if (r == 0) {
      i++;
      continue;
    }
    e->inflight[i] = e->inflight.back();
    e->inflight.pop_back();
    if (r < 0) bytes = 0;
    if (f.is_read) complete_read(f.ep, (size_t)bytes, f.rlen);
    else complete_write(f.ep, bytes);
  }
  if (e->q_reads.empty() && e->q_writes.empty()) return true;
  std::vector<b200_endpoint*> reads, writes;
  reads.swap(e->q_reads);
  writes.swap(e->q_writes);
  bool available = true;
  for (size_t k = 0; k < writes.size(); k++) {
    b200_endpoint* ep = writes[k];
    if (!ep->write_cb) {
      ep->queued_write = false;
      continue;
    }
    int again = 0;
    void* op = available ? ops->post_send(ep->pair, ep->outgoing + ep->outgoing_idx, ep->outgoing_count - ep->outgoing_idx,
                                          ep->outgoing_byte_idx, B200_BATCH_UNTIL_BLOCKED, &again)
                         : nullptr;
    if (op) {
      ep->queued_write = false;
      e->inflight.push_back({ep, false, op, 0});
      e->bstats[1]++;
    } else if (!available || again) {
      if (again == 2) available = false;
      e->q_writes.push_back(ep);  // stays queued
    } else {
      ep->queued_write = false;
      complete_write(ep, 0);
    }
  }
  for (size_t k = 0; k < reads.size(); k++) {
    b200_endpoint* ep = reads[k];
    if (!ep->read_cb) {
      ep->queued_read = false;
      continue;
    }
    int again = 0;
    void* op = nullptr;
    if (available) {
      prepare_read_slice(ep);
      ep->inq = 1;
      op = ops->post_recv(ep->pair, ep->incoming.empty() ? nullptr : ep->incoming[0].ptr(),
                          ep->incoming.empty() ? 0 : ep->incoming[0].len, B200_BATCH_UNTIL_BLOCKED, &again);
    }
    if (op) {
      ep->queued_read = false;
      e->inflight.push_back({ep, true, op, buf_length(ep->incoming)});
      e->bstats[2]++;
    } else if (!available || again) {
      if (again == 2) available = false;
      e->q_reads.push_back(ep);
    } else {
      ep->queued_read = false;
      complete_read(ep, 0, buf_length(ep->incoming));
    }
  }
  if (available) e->bstats[0]++;
  return available;
}

}  // namespace

// ============================================================ public: engine

extern "C" b200_engine* b200_engine_create(const b200_pair_ops* ops, int busy_poll_us) {
  b200_engine* e = new b200_engine();
  e->ops = ops ? ops : &kCudaOps;
  e->epfd = epoll_create1(EPOLL_CLOEXEC);
  e->wake_fd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  struct epoll_event ev;
  ev.events = (uint32_t)EPOLLIN;
  ev.data.ptr = reinterpret_cast<void*>(kTagWake);
  epoll_ctl(e->epfd, EPOLL_CTL_ADD, e->wake_fd, &ev);
  if (busy_poll_us < 0) {
    const char* v = getenv("GRPC_RDMA_BUSY_POLLING_TIMEOUT_US");  // config.cc:75-81
    busy_poll_us = v ? atoi(v) : 500;
  }
  e->busy_poll_us = busy_poll_us;
  const char* b = getenv("B200_ENDPOINT_BATCH");
  e->batch = e->ops->submit != nullptr && (!b || atoi(b) != 0);
  const char* a = getenv("B200_ENDPOINT_ASYNC");
  // off by default: measured slower than the pass-at-a-time form on the streaming workload (reads get posted as
  // soon as the first frames land and come back small; profiles/r2_endpoint_stream.md)
  e->async = e->batch && e->ops->post_send && e->ops->post_recv && e->ops->poll && a && atoi(a) != 0;
  const char* c = getenv("B200_ENDPOINT_READ_CHUNK_KB");
  if (c && atol(c) > 0) e->max_read_chunk = (size_t)atol(c) * 1024;
  e->pool = std::make_shared<Pool>(e->ops);
  return e;
}

extern "C" void b200_engine_destroy(b200_engine* e) {
  if (!e) return;
  if (e->epfd >= 0) close(e->epfd);
  if (e->wake_fd >= 0) close(e->wake_fd);
  delete e;
}

extern "C" void b200_engine_stats(b200_engine* e, uint64_t out[4]) {
  Lock lk(e->mu);
  for (int i = 0; i < 4; i++) out[i] = e->stats[i];
}
extern "C" void b200_engine_batch_stats(b200_engine* e, uint64_t out[3]) {
  Lock lk(e->mu);
  for (int i = 0; i < 3; i++) out[i] = e->bstats[i];
}

// pollable_epoll (:1079-1176) + pollable_process_events (:977-1066).  Like the reference (rdma_mu, :1103-1145)
// the lock is held around one scan of the pairs only: not between scans, not across epoll_wait, not while
// the GPU works on a submit and not while a user callback runs.
extern "C" int b200_engine_work(b200_engine* e, int timeout_ms) {
  const b200_pair_ops* ops = e->ops;
  struct Ev {
    b200_endpoint* ep;
    uint32_t events;
    bool tagged;
  };
  std::vector<Ev> evs;
  int64_t polling_timeout_us = e->busy_poll_us;
  if (timeout_ms >= 0) polling_timeout_us = std::min<int64_t>((int64_t)timeout_ms * 1000, polling_timeout_us);
  const auto begin = std::chrono::steady_clock::now();
  int64_t elapsed_us = 0;
  bool queued = false;
  do {  // busy-poll window over the pairs (:1104-1145)
    {
      Lock lk(e->mu);
      for (size_t i = 0; i < e->rdma_fds.size() && evs.size() < (size_t)kMaxEpollEvents; i++) {
        b200_endpoint* ep = e->rdma_fds[i];
        const int status = ops->status(ep->pair);
        if (status == B200_CONNECTED) {
          uint32_t events = 0;
          if (ops->has_message(ep->pair)) events |= EPOLLIN;
          if (ops->has_pending_writes(ep->pair)) events |= EPOLLOUT;
          if (events) evs.push_back({ep, events, false});
        } else if (status == B200_HALF_CLOSED || status == B200_ERROR) {
          evs.push_back({ep, EPOLLIN, false});  // so that do_read handles the close
        }
      }
      queued = !e->q_reads.empty() || !e->q_writes.empty() || !e->exec.empty() || !e->inflight.empty();
    }
    elapsed_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - begin).count();
  } while (evs.empty() && !queued && elapsed_us < polling_timeout_us);
  if (!evs.empty()) {
    Lock lk(e->mu);
    e->stats[0]++;
    e->stats[2] += evs.size();
  } else if (!queued) {  // busy-polling timed out: switch to epoll on the pairs' eventfds (:1151-1163)
    int timeout = timeout_ms;
    if (timeout > 0) timeout = (int)std::max<int64_t>(0, timeout - elapsed_us / 1000);
    struct epoll_event eev[kMaxEpollEvents];
    int r;
    do {
      r = epoll_wait(e->epfd, eev, kMaxEpollEvents, timeout);
    } while (r < 0 && errno == EINTR);
    if (r < 0) return -1;
    Lock lk(e->mu);
    e->stats[1]++;
    for (int i = 0; i < r; i++) {
      const intptr_t tag = reinterpret_cast<intptr_t>(eev[i].data.ptr);
      if (tag == kTagWake) {
        eventfd_t v;
        (void)eventfd_read(e->wake_fd, &v);
      } else if (tag & kTagPair) {
        evs.push_back({reinterpret_cast<b200_endpoint*>(tag & ~kTagPair), eev[i].events, true});
      }
    }
    e->stats[3] += evs.size();
  }
  Lock lk(e->mu);
  e->in_work = true;
  for (const Ev& ev : evs) {
    b200_endpoint* ep = ev.ep;
    if (std::find(e->rdma_fds.begin(), e->rdma_fds.end(), ep) == e->rdma_fds.end()) continue;  // orphaned meanwhile
    if (ev.tagged) {  // eventfd kicked by the background poller (:1010-1035)
      const int status = ops->status(ep->pair);
      if (status != B200_UNINITIALIZED && status != B200_DISCONNECTED) {
        ops->consume_wakeup(ep->pair);
        if (ops->has_message(ep->pair) || ops->status(ep->pair) == B200_HALF_CLOSED) set_ready(ep, ep->rd, true);
        if (ops->has_pending_writes(ep->pair)) set_ready(ep, ep->wr, false);
      }
    } else {  // synthetic events of the scan (:1036-1066)
      if (ev.events & (EPOLLIN | EPOLLPRI | EPOLLHUP)) set_ready(ep, ep->rd, true);
      if (ev.events & (EPOLLOUT | EPOLLHUP)) set_ready(ep, ep->wr, false);
    }
  }
  flush(e, lk);  // closures scheduled meanwhile (errors, call-by-call continuations)
  if (e->batch) {
    eventfd_t v;
    (void)eventfd_read(e->wake_fd, &v);
    if (!e->async || !pump(e)) run_batch(e, lk);
    flush(e, lk);
  }
  e->in_work = false;
  if (!e->q_reads.empty() || !e->q_writes.empty()) kick_engine(e);  // queued by the callbacks: next pass must not sleep
  return (int)evs.size();
}

// ========================================================== public: endpoint

extern "C" int b200_exchange_data(int fd, const char* buf_in, char* buf_out, size_t sz) {
  size_t bytes_send = 0, bytes_recv = 0;
  if (fd < 3) return -1;  // rdma_bp_posix.cc:642-648
  struct pollfd fds[1];
  fds[0].fd = fd;
  while (bytes_recv < sz) {
    fds[0].events = POLLIN | (bytes_send < sz ? POLLOUT : 0);
    const int r = poll(fds, 1, -1);
    if (r <= 0) continue;
    if (bytes_send < sz && (fds[0].revents & POLLOUT)) {
      const ssize_t n = ::write(fd, buf_in + bytes_send, sz - bytes_send);
      if (n < 0) {
        if (errno != EINTR && errno != EAGAIN) return -1;
      } else {
        bytes_send += (size_t)n;
      }
    }
    if (fds[0].revents & POLLIN) {
      const ssize_t n = ::read(fd, buf_out + bytes_recv, sz - bytes_recv);
      if (n < 0) {
        if (errno != EINTR && errno != EAGAIN) return -1;
      } else if (n == 0) {
        return -1;  // peer went away during bootstrap
      } else {
        bytes_recv += (size_t)n;
      }
    }
    if (fds[0].revents & (POLLERR | POLLNVAL)) return -1;
  }
  while (bytes_send < sz) {  // our half may still be in flight
    const ssize_t n = ::write(fd, buf_in + bytes_send, sz - bytes_send);
    if (n < 0) {
      if (errno != EINTR && errno != EAGAIN) return -1;
    } else {
      bytes_send += (size_t)n;
    }
  }
  return 0;
}

extern "C" b200_endpoint* b200_endpoint_create(b200_engine* e, int fd, const char* peer_string, int enable_poller) {
  const b200_pair_ops* ops = e->ops;
  b200_endpoint* ep = new b200_endpoint();
  ep->engine = e;
  ep->fd = fd;
  ep->peer_string = peer_string ? peer_string : "";
  void* pair = ops->pool_take(ep->peer_string.c_str());  // :761
  if (!pair) {
    delete ep;
    return nullptr;
  }
  ops->init(pair);  // :765
  char self[B200_ADDRESS_BYTES], peer[B200_ADDRESS_BYTES];
  const size_t n = ops->self_address(pair, self);
  // the bootstrap blocks on the accept/connect thread exactly like the reference (:770-773)
  if (n != B200_ADDRESS_BYTES || b200_exchange_data(fd, self, peer, n) != 0 || !ops->connect(pair, peer, n)) {
    ops->disconnect(pair);  // cleanup, :778-783
    ops->pool_putback(pair);
    delete ep;
    return nullptr;
  }
  ep->pair = pair;
  ep->enable_poller = enable_poller != 0;
  Lock lk(e->mu);
  // grpc_fd_set_arg + pollable_add_fd: the pair's eventfd joins the epoll set with tag ptr|2
  struct epoll_event ev;
  ev.events = (uint32_t)(EPOLLIN | EPOLLET);
  ev.data.ptr = reinterpret_cast<void*>(reinterpret_cast<intptr_t>(ep) | kTagPair);
  epoll_ctl(e->epfd, EPOLL_CTL_ADD, ops->wakeup_read_fd(pair), &ev);
  e->rdma_fds.push_back(ep);
  if (ep->enable_poller) ops->poller_add(pair);  // :790-793
  return ep;
}

// rdma_read, rdma_bp_posix.cc:343-375
extern "C" void b200_endpoint_read(b200_endpoint* ep, b200_closure_fn cb, void* arg, int urgent) {
  b200_engine* e = ep->engine;
  Lock lk(e->mu);
  if (ep->read_cb != nullptr) abort();  // GPR_ASSERT(rdma->read_cb == nullptr)
  ep->read_cb = cb;
  ep->read_arg = arg;
  ep->incoming.clear();
  ep->incoming.swap(ep->last_read_buffer);  // reuse what the last read did not fill
  ep->refs++;
  if (ep->is_first_read) {
    ep->is_first_read = false;
    notify_on(ep, ep->rd, true);
  } else if (!urgent && ep->inq == 0) {
    notify_on(ep, ep->rd, true);
  } else {
    dispatch(ep, true);
  }
  flush(e, lk);
}

extern "C" size_t b200_endpoint_incoming(b200_endpoint* ep, const b200_slice** slices) {
  if (slices) *slices = ep->incoming_view.data();
  return ep->incoming_view.size();
}

// rdma_write, rdma_bp_posix.cc:559-588
extern "C" void b200_endpoint_write(b200_endpoint* ep, const b200_slice* slices, size_t n, b200_closure_fn cb,
                                    void* arg) {
  b200_engine* e = ep->engine;
  Lock lk(e->mu);
  if (ep->write_cb != nullptr) abort();  // GPR_ASSERT(rdma->write_cb == nullptr)
  uint64_t length = 0;
  for (size_t i = 0; i < n; i++) length += slices[i].len;
  if (length == 0) {  // :566-574
    if (ep->shutdown) schedule_user(e, cb, arg, annotate(ep, "EOF").c_str());
    else schedule_user(e, cb, arg, nullptr);
    flush(e, lk);
    return;
  }
  ep->outgoing = slices;
  ep->outgoing_count = n;
  ep->outgoing_idx = 0;
  ep->outgoing_byte_idx = 0;
  if (e->batch) {  // the flush loop runs with everybody else's in the next engine pass
    ep->refs++;
    ep->write_cb = cb;
    ep->write_arg = arg;
    dispatch(ep, false);
    flush(e, lk);
    return;
  }
  std::string err;
  if (!rdma_flush(ep, &err)) {
    ep->refs++;
    ep->write_cb = cb;
    ep->write_arg = arg;
    notify_on(ep, ep->wr, false);
  } else {
    schedule_user(e, cb, arg, err.empty() ? nullptr : err.c_str());
  }
  flush(e, lk);
}

// rdma_shutdown, rdma_bp_posix.cc:100-104 (grpc_fd_shutdown)
extern "C" void b200_endpoint_shutdown(b200_endpoint* ep, const char* why) {
  b200_engine* e = ep->engine;
  Lock lk(e->mu);
  if (ep->shutdown) return;
  ep->shutdown = true;
  ep->shutdown_why = annotate(ep, why ? why : "endpoint shutdown");
  std::string es = ep->shutdown_why;
  if (ep->rd.armed) {
    ep->rd.armed = false;
    schedule(e, [ep, es] { rdma_handle_read(ep, es.c_str()); });
  }
  if (ep->wr.armed) {
    ep->wr.armed = false;
    schedule(e, [ep, es] { rdma_handle_write(ep, es.c_str()); });
  }
  flush(e, lk);
}

// rdma_destroy, rdma_bp_posix.cc:168 -> RDMA_UNREF
extern "C" void b200_endpoint_destroy(b200_endpoint* ep) {
  b200_engine* e = ep->engine;
  Lock lk(e->mu);
  ep->last_read_buffer.clear();
  ep_unref(ep);
}

extern "C" const char* b200_endpoint_peer(b200_endpoint* ep) { return ep->peer_string.c_str(); }
extern "C" int b200_endpoint_fd(b200_endpoint* ep) { return ep->fd; }
extern "C" void* b200_endpoint_pair(b200_endpoint* ep) { return ep->pair; }
