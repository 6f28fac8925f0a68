// pingpong.cc -- BENCH DRIVER (not product code): unary ping-pong over the C ABI of
// include/b200_pair.h, the B200 counterpart of oracle/ref_harness.cc:ref_bench_pingpong
// (BASELINE config 3: M-byte request, M-byte echo, 1..N connections).
//
// `groups` client threads and `groups` server threads; client i and server i own the same
// conns/groups connections.  A client sends on a connection, spins on HasMessage (a wait-free read
// of the pair's host-visible mirror), receives the echo and moves to its next connection; a server
// polls its connections round-robin, receives and echoes.  Buffers are registered (pinned) host
// memory: every request and every echo crosses PCIe in both directions.  With the service kernel
// running (b200_service_start) no call launches anything.
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/b200_pair.h"

namespace {
struct Worker {
  int first, n, iters, warm, is_server;
  uint64_t msg;
  b200_pair** cli;
  b200_pair** srv;
  uint8_t* buf;  // this thread's registered buffer: [0,msg) out, [msg, 2 msg) in
  uint64_t* rtt_ns;
  pthread_barrier_t* bar;
  std::atomic<int>* stop;
  int err;
  uint64_t t_send = 0, t_wait = 0, t_recv = 0, n_ops = 0;  // phase times (ns), B200_PP_TRACE
};

inline uint64_t now_ns() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

void* run(void* arg) {
  Worker* w = (Worker*)arg;
  for (uint64_t i = 0; i < w->msg; i++) w->buf[i] = (uint8_t)(i * 7 + w->first);
  pthread_barrier_wait(w->bar);
  if (w->is_server) {
    std::vector<uint64_t> got(w->n, 0);
    while (!w->stop->load(std::memory_order_relaxed)) {
      for (int c = 0; c < w->n; c++) {
        b200_pair* p = w->srv[w->first + c];
        if (!b200_pair_has_message(p)) continue;
        const uint64_t s0 = now_ns();
        got[c] += b200_pair_recv(p, w->buf + w->msg + got[c], w->msg - got[c]);
        const uint64_t s1 = now_ns();
        w->t_recv += s1 - s0;
        if (got[c] == w->msg) {
          got[c] = 0;
          b200_slice sl{w->buf + w->msg, w->msg};
          uint64_t sent = 0;
          while (sent < w->msg && !w->stop->load(std::memory_order_relaxed)) sent += b200_pair_send(p, &sl, 1, sent);
          w->t_send += now_ns() - s1;
          w->n_ops++;
        }
      }
    }
  } else {
    for (int k = -w->warm; k < w->iters && !w->err; k++) {
      for (int c = 0; c < w->n; c++) {
        b200_pair* p = w->cli[w->first + c];
        w->buf[0] = (uint8_t)k;
        const uint64_t t0 = now_ns();
        b200_slice sl{w->buf, w->msg};
        uint64_t sent = 0;
        while (sent < w->msg) sent += b200_pair_send(p, &sl, 1, sent);
        const uint64_t t1 = now_ns();
        uint64_t got = 0, twait = 0, trecv = 0;
        while (got < w->msg) {
          const uint64_t a0 = now_ns();
          while (!b200_pair_has_message(p)) {
            if (now_ns() - t0 > 20000000000ull) {
              w->err = 1;
              return nullptr;
            }
          }
          const uint64_t a1 = now_ns();
          got += b200_pair_recv(p, w->buf + w->msg + got, w->msg - got);
          twait += a1 - a0;
          trecv += now_ns() - a1;
        }
        if (k >= 0) {
          w->rtt_ns[(size_t)(w->first + c) * w->iters + k] = now_ns() - t0;
          w->t_send += t1 - t0;
          w->t_wait += twait;
          w->t_recv += trecv;
          w->n_ops++;
        }
        if (memcmp(w->buf, w->buf + w->msg, w->msg) != 0) w->err = 2;  // echo must be bit-exact
      }
    }
  }
  return nullptr;
}
}  // namespace

// Returns wall seconds of the timed part (< 0 on failure: -1 setup, -2 timeout, -3 corrupt echo).
extern "C" double b200_pp_run(int conns, int groups, int iters, int warm, uint64_t msg_bytes, uint64_t* rtt_ns) {
  if (groups < 1) groups = 1;
  if (groups > conns) groups = conns;
  std::vector<b200_pair*> cli(conns), srv(conns);
  for (int c = 0; c < conns; c++) {
    cli[c] = b200_pool_take(("pp-cli-" + std::to_string(c)).c_str());
    srv[c] = b200_pool_take(("pp-srv-" + std::to_string(c)).c_str());
    if (!cli[c] || !srv[c]) return -1;
    b200_pair_init(cli[c]);
    b200_pair_init(srv[c]);
    char a[B200_ADDRESS_BYTES], b[B200_ADDRESS_BYTES];
    b200_pair_self_address(cli[c], a);
    b200_pair_self_address(srv[c], b);
    if (!b200_pair_connect(cli[c], b, sizeof b) || !b200_pair_connect(srv[c], a, sizeof a)) return -1;
  }
  std::atomic<int> stop{0};
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, 2 * groups + 1);
  std::vector<pthread_t> th(2 * groups);
  std::vector<Worker> ws(2 * groups);
  std::vector<void*> bufs(2 * groups);
  int base = 0;
  for (int g = 0; g < groups; g++) {
    const int nc = conns / groups + (g < conns % groups ? 1 : 0);
    for (int s = 0; s < 2; s++) {
      bufs[2 * g + s] = b200_mem_alloc_host(2 * msg_bytes + 64);
      if (!bufs[2 * g + s]) return -1;
      ws[2 * g + s] = Worker{base, nc, iters, warm, s, msg_bytes, cli.data(), srv.data(), (uint8_t*)bufs[2 * g + s],
                             rtt_ns, &bar, &stop, 0};
    }
    base += nc;
  }
  for (int i = 0; i < 2 * groups; i++) pthread_create(&th[i], nullptr, run, &ws[i]);
  pthread_barrier_wait(&bar);
  const uint64_t t0 = now_ns();
  for (int g = 0; g < groups; g++) pthread_join(th[2 * g], nullptr);
  const uint64_t t1 = now_ns();
  stop.store(1);
  for (int g = 0; g < groups; g++) pthread_join(th[2 * g + 1], nullptr);
  pthread_barrier_destroy(&bar);
  int err = 0;
  for (auto& w : ws) err |= w.err;
  if (getenv("B200_PP_TRACE")) {
    uint64_t cs = 0, cw = 0, cr = 0, cn = 0, ss = 0, sr = 0, sn = 0;
    for (auto& w : ws) {
      if (w.is_server) { ss += w.t_send; sr += w.t_recv; sn += w.n_ops; }
      else { cs += w.t_send; cw += w.t_wait; cr += w.t_recv; cn += w.n_ops; }
    }
    if (cn && sn)
      fprintf(stderr, "pp trace conns=%d: client send %.2f us, wait-for-echo %.2f us, recv %.2f us | server recv %.2f us, send %.2f us\n",
              conns, cs / 1e3 / cn, cw / 1e3 / cn, cr / 1e3 / cn, sr / 1e3 / sn, ss / 1e3 / sn);
  }
  for (int c = 0; c < conns; c++) {
    b200_pair_disconnect(cli[c]);
    b200_pair_disconnect(srv[c]);
    b200_pool_putback(cli[c]);
    b200_pool_putback(srv[c]);
  }
  for (void* b : bufs) b200_mem_free_host(b);
  if (err & 1) return -2;
  if (err & 2) return -3;
  return 1e-9 * (double)(t1 - t0);
}
