// ep_stream.cc -- BENCH DRIVER (not product code): streaming through the drop-in surface.
//
// BASELINE configs[1] driven the way chttp2 drives a grpc_endpoint: `conns` connections, each a pair of
// endpoints (include/b200_endpoint.h = rdma_bp_posix.cc + the BPEV poll loop) created over a socketpair
// bootstrap; a client thread keeps ONE message per connection in flight with b200_endpoint_write (the message
// cut into the slice list chttp2 produces: 9-byte DATA frame headers + <= 16384-byte payload slices, every
// slice at its own address with a gap to the next one: nothing is adjacent, nothing can be coalesced), a
// server thread reads with b200_endpoint_read and compares EVERY delivered byte with what was sent; both spin
// b200_engine_work.  `threads` client/server thread pairs share the connections.  The same code runs over any
// pair-ops table: NULL = the CUDA library (batching engine, service kernels), or the reference's own
// PairPollable (tests/native/ref_pair_ops.cc) for the CPU number beside it.
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200_endpoint.h"

namespace {

inline uint64_t now_ns() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

struct Shared {
  const b200_pair_ops* ops;
  int msgs, warm;
  uint64_t msg_bytes, stream_bytes;  // payload of one message; bytes the endpoint moves for it (payload + framing)
  std::vector<uint64_t> lens;        // the slice list of one message
  std::atomic<int> ready{0}, go{0}, err{0};
  std::atomic<uint64_t> t_start{0}, t_end{0};
  std::atomic<int> servers_done{0}, warm_done{0};
  int nthreads;
  int verify = 1;  // B200_EP_VERIFY=0: experiment only (where does the time go?) -- results are then not reported
};

struct Conn {
  b200_endpoint* cep = nullptr;
  b200_endpoint* sep = nullptr;
  // source: two different messages (A/B), each as (a) the contiguous expected stream and (b) the slice list
  // carved out of a registered arena with gaps
  uint8_t* expect[2] = {nullptr, nullptr};
  uint8_t* arena[2] = {nullptr, nullptr};
  std::vector<b200_slice> slices[2];
  // client state
  int sent = 0, write_busy = 0;
  // server state
  uint64_t got = 0;  // bytes of the current message received so far
  int received = 0;  // complete messages
  int reading = 0;
  struct Side* cside = nullptr;
  struct Side* sside = nullptr;
};

struct Side {  // one thread
  Shared* sh;
  b200_engine* eng;
  std::vector<Conn*> conns;
  int is_server;
  int total_msgs;
  uint64_t bad = 0;
};

void on_write(void* arg, const char* error) {
  Conn* c = (Conn*)arg;
  c->write_busy = 0;
  if (error) c->cside->sh->err = 1;
  else c->sent++;
}

void on_read(void* arg, const char* error) {
  Conn* c = (Conn*)arg;
  Side* s = c->sside;
  c->reading = 0;
  if (error) {
    s->sh->err = 2;
    return;
  }
  const b200_slice* in;
  const size_t n = b200_endpoint_incoming(c->sep, &in);
  for (size_t i = 0; i < n; i++) {
    const uint8_t* p = (const uint8_t*)in[i].ptr;
    uint64_t len = in[i].len;
    while (len) {  // a read may straddle two messages
      const uint64_t left = s->sh->stream_bytes - c->got;
      const uint64_t take = len < left ? len : left;
      if (s->sh->verify && memcmp(p, c->expect[c->received & 1] + c->got, take) != 0) s->bad++;
      c->got += take;
      p += take;
      len -= take;
      if (c->got == s->sh->stream_bytes) {
        c->got = 0;
        c->received++;
      }
    }
  }
}

void* side_main(void* arg) {
  Side* s = (Side*)arg;
  Shared* sh = s->sh;
  sh->ready++;
  while (!sh->go.load()) {
  }
  const int total = sh->warm + sh->msgs;
  bool warm_reported = false, finished = false;
  const uint64_t deadline = now_ns() + 120ull * 1000000000ull;
  while (!finished && !sh->err.load()) {
    int done = 0, warm = 0;
    for (Conn* c : s->conns) {
      if (s->is_server) {
        if (c->received >= total) {
          done++;
          warm++;
          continue;
        }
        if (c->received >= sh->warm) warm++;
        if (!c->reading) {
          c->reading = 1;
          b200_endpoint_read(c->sep, on_read, c, 0);
        }
      } else {
        if (c->sent >= total) {
          done++;
          warm++;
          continue;
        }
        if (c->sent >= sh->warm) warm++;
        // the timed messages start only once every connection of the job has finished its warm-up
        if (!c->write_busy && (c->sent < sh->warm || sh->t_start.load() != 0)) {
          c->write_busy = 1;
          const int par = c->sent & 1;
          b200_endpoint_write(c->cep, c->slices[par].data(), c->slices[par].size(), on_write, c);
        }
      }
    }
    if (s->is_server && !warm_reported && warm == (int)s->conns.size()) {
      warm_reported = true;
      if (++sh->warm_done == sh->nthreads) sh->t_start = now_ns();
    }
    finished = done == (int)s->conns.size();
    if (!finished) b200_engine_work(s->eng, 0);
    if (now_ns() > deadline) sh->err = 3;
  }
  if (s->is_server) {
    if (++sh->servers_done == sh->nthreads) sh->t_end = now_ns();
  } else {
    // keep the engine turning until the servers have everything (partial writes are driven from here)
    while (sh->servers_done.load() < sh->nthreads && !sh->err.load()) b200_engine_work(s->eng, 0);
  }
  return nullptr;
}

}  // namespace

// Returns seconds of the timed part (all servers warm -> all servers have every message); < 0 on failure.
// out[0] = payload bytes delivered in the timed part, out[1] = bytes that differed (must be 0),
// out[2] = submit calls of the client engines, out[3] = submit calls of the server engines.
extern "C" double ep_stream_run(const b200_pair_ops* ops, int conns, int threads, int msgs, int warm, uint64_t msg_bytes,
                                int busy_us, uint64_t* out) {
  if (threads < 1) threads = 1;
  if (threads > conns) threads = conns;
  Shared sh;
  sh.ops = ops;
  sh.msgs = msgs;
  sh.warm = warm;
  sh.msg_bytes = msg_bytes;
  sh.nthreads = threads;
  if (getenv("B200_EP_VERIFY") && atoi(getenv("B200_EP_VERIFY")) == 0) sh.verify = 0;
  {  // chttp2_slice_lens (grpc-rdma_b200/__init__.py): 5-byte gRPC prefix, 16384-byte DATA frames
    uint64_t data = 5 + msg_bytes;
    while (data > 0) {
      const uint64_t n = data < 16384 ? data : 16384;
      sh.lens.push_back(9);
      sh.lens.push_back(n);
      data -= n;
    }
  }
  sh.stream_bytes = 0;
  for (uint64_t l : sh.lens) sh.stream_bytes += l;
  std::vector<Conn> cs(conns);
  std::vector<Side> sides(2 * threads);
  std::vector<b200_engine*> engines(2 * threads);
  for (int t = 0; t < 2 * threads; t++) {
    engines[t] = b200_engine_create(ops, busy_us);
    sides[t].sh = &sh;
    sides[t].eng = engines[t];
    sides[t].is_server = t & 1;
  }
  auto alloc = [&](size_t n) -> uint8_t* {
    void* p = (ops && ops->mem_alloc) ? ops->mem_alloc(n) : (!ops ? b200_mem_alloc_host(n) : malloc(n));
    return (uint8_t*)p;
  };
  int rc = 0;
  for (int c = 0; c < conns && !rc; c++) {
    Conn& k = cs[c];
    const int t = c % threads;
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) {
      rc = -1;
      break;
    }
    std::thread th([&] { k.sep = b200_endpoint_create(engines[2 * t + 1], sv[1], "ipv4:server", 0); });
    k.cep = b200_endpoint_create(engines[2 * t], sv[0], "ipv4:client", 0);
    th.join();
    if (!k.cep || !k.sep) {
      rc = -2;
      break;
    }
    k.cside = &sides[2 * t];
    k.sside = &sides[2 * t + 1];
    sides[2 * t].conns.push_back(&k);
    sides[2 * t + 1].conns.push_back(&k);
    for (int par = 0; par < 2; par++) {
      k.expect[par] = (uint8_t*)malloc(sh.stream_bytes);
      const size_t arena_bytes = sh.stream_bytes + sh.lens.size() * 256 + 4096;
      k.arena[par] = alloc(arena_bytes);
      if (!k.expect[par] || !k.arena[par]) {
        rc = -3;
        break;
      }
      uint64_t x = 0x9E3779B97F4A7C15ull ^ ((uint64_t)c << 32 | (uint64_t)par);
      uint64_t* w = (uint64_t*)k.expect[par];
      for (uint64_t i = 0; i < sh.stream_bytes / 8; i++) {  // xorshift64*: incompressible, differs per (conn, parity)
        x ^= x >> 12;
        x ^= x << 25;
        x ^= x >> 27;
        w[i] = x * 0x2545F4914F6CDD1Dull;
      }
      for (uint64_t i = sh.stream_bytes & ~7ull; i < sh.stream_bytes; i++) k.expect[par][i] = (uint8_t)(i * 131 + c);
      size_t off = 0, pos = 0;
      uint64_t g = (uint64_t)c * 7 + par;
      for (uint64_t l : sh.lens) {
        g = g * 6364136223846793005ull + 1442695040888963407ull;
        off += 8 + (g >> 58) * 3;  // a gap of 8..197 bytes: every slice at its own, differently aligned address
        memcpy(k.arena[par] + off, k.expect[par] + pos, l);
        k.slices[par].push_back({k.arena[par] + off, l});
        off += l;
        pos += l;
      }
    }
  }
  std::vector<pthread_t> th(2 * threads);
  if (!rc) {
    for (int t = 0; t < 2 * threads; t++) pthread_create(&th[t], nullptr, side_main, &sides[t]);
    while (sh.ready.load() < 2 * threads) {
    }
    sh.go = 1;
    for (int t = 0; t < 2 * threads; t++) pthread_join(th[t], nullptr);
  }
  uint64_t bad = 0, csub = 0, ssub = 0;
  for (int t = 0; t < 2 * threads; t++) {
    bad += sides[t].bad;
    uint64_t b[3] = {0, 0, 0};
    b200_engine_batch_stats(engines[t], b);
    (t & 1 ? ssub : csub) += b[0];
  }
  for (Conn& k : cs) {
    if (k.cep) b200_endpoint_destroy(k.cep);
    if (k.sep) b200_endpoint_destroy(k.sep);
  }
  for (b200_engine* e : engines) b200_engine_destroy(e);
  for (Conn& k : cs)
    for (int par = 0; par < 2; par++) {
      free(k.expect[par]);
      if (k.arena[par]) {
        if (ops && ops->mem_free) ops->mem_free(k.arena[par]);
        else if (!ops) b200_mem_free_host(k.arena[par]);
        else free(k.arena[par]);
      }
    }
  if (out) {
    out[0] = (uint64_t)conns * (uint64_t)msgs * msg_bytes;
    out[1] = bad;
    out[2] = csub;
    out[3] = ssub;
  }
  if (rc) return rc;
  if (sh.err.load()) return -10 - sh.err.load();
  if (bad) return -20;
  return 1e-9 * (double)(sh.t_end.load() - sh.t_start.load());
}
