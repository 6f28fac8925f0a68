// endpoint_driver.cc -- TEST INFRASTRUCTURE: drives the product's endpoint + BPEV poll loop
// (include/b200_endpoint.h) the way the reference's own tests drive a grpc_endpoint.
//
//   drv_read_and_write   test/core/iomgr/endpoint_tests.cc:read_and_write_test (:215-300): one endpoint
//                        writes a byte ramp b[i] = i mod 256 in writes of `write_size` bytes made of
//                        `slice_size`-byte slices, the other reads and checks the ramp; optional
//                        shutdown of both ends right after the first read/write are started.
//   drv_shutdown_sequence  endpoint_tests.cc:multiple_shutdown_test (:325-360)
//   drv_echo             examples/cpp/test/{common.h:5-32,greeter_client.cc:44-67}: N random
//                        messages of 1 .. max_len bytes echoed by a server, client checks
//                        msg == reply; client and server each run their own engine on their own thread.
//   drv_peer_close       rdma_do_read's HalfClosed branch (rdma_bp_posix.cc:218-228): destroying one
//                        end makes the other end's pending read fail with "Pair closed".
//
// ops == NULL drives the CUDA library (GPU tests); the CPU tests pass the oracle table
// (tests/native/oracle_pair_ops.c).  Every function returns 0 on success, a line number on failure.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200_endpoint.h"

namespace {

#define FAIL() return __LINE__
using Clock = std::chrono::steady_clock;

struct Fixture {
  b200_engine* eng[2] = {nullptr, nullptr};
  b200_endpoint* ep[2] = {nullptr, nullptr};
};

// grpc_endpoint_test_fixture over a socketpair (tcp_posix_test.cc:590-613); the two creates block on
// exchange_data against each other, like a connecting client and an accepting server.
bool make_fixture(const b200_pair_ops* ops, int busy_us, int enable_poller, bool two_engines, Fixture* f) {
  int sv[2];
  if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) return false;
  // descriptors 0-2 are refused by exchange_data (rdma_bp_posix.cc:642-648)
  f->eng[0] = b200_engine_create(ops, busy_us);
  f->eng[1] = two_engines ? b200_engine_create(ops, busy_us) : f->eng[0];
  std::thread t([&] { f->ep[1] = b200_endpoint_create(f->eng[1], sv[1], "ipv4:server", enable_poller); });
  f->ep[0] = b200_endpoint_create(f->eng[0], sv[0], "ipv4:client", enable_poller);
  t.join();
  return f->ep[0] && f->ep[1];
}

void drop_fixture(Fixture* f) {
  if (f->eng[1] != f->eng[0]) b200_engine_destroy(f->eng[1]);
  b200_engine_destroy(f->eng[0]);
}

struct RW {
  b200_endpoint *read_ep, *write_ep;
  uint64_t target, bytes_read = 0, bytes_written = 0, cur_write, slice_size;
  int cur_read_data = 0;
  uint8_t cur_write_data = 0;
  int read_done = 0, write_done = 0, bad = 0;
  std::vector<std::vector<uint8_t>> blocks;
  std::vector<b200_slice> sl;
  std::string read_err, write_err;
};

void rw_read_cb(void* arg, const char* error);
void rw_write_cb(void* arg, const char* error);

void rw_read_cb(void* arg, const char* error) {
  RW* s = (RW*)arg;
  if (!error) {  // count_slices, endpoint_tests.cc:52-66
    const b200_slice* in;
    size_t n = b200_endpoint_incoming(s->read_ep, &in);
    for (size_t i = 0; i < n; i++) {
      const uint8_t* p = (const uint8_t*)in[i].ptr;
      for (uint64_t j = 0; j < in[i].len; j++) {
        if (p[j] != s->cur_read_data) s->bad++;
        s->cur_read_data = (s->cur_read_data + 1) % 256;
      }
      s->bytes_read += in[i].len;
    }
    if (s->bytes_read == s->target) {
      s->read_done = 1;
    } else {
      b200_endpoint_read(s->read_ep, rw_read_cb, s, 0);
    }
  } else {
    s->read_err = error;
    s->read_done = 2;
  }
}

void rw_write_cb(void* arg, const char* error) {
  RW* s = (RW*)arg;
  if (!error) {
    s->bytes_written += s->cur_write;
    if (s->target - s->bytes_written < s->cur_write) s->cur_write = s->target - s->bytes_written;
    if (s->cur_write != 0) {  // allocate_blocks, endpoint_tests.cc:77-99
      s->blocks.clear();
      s->sl.clear();
      uint64_t left = s->cur_write;
      while (left) {
        uint64_t n = s->slice_size > left ? left : s->slice_size;
        std::vector<uint8_t> b(n);
        for (uint64_t j = 0; j < n; j++) b[j] = s->cur_write_data++;
        s->blocks.push_back(std::move(b));
        left -= n;
      }
      for (auto& b : s->blocks) s->sl.push_back({b.data(), b.size()});
      b200_endpoint_write(s->write_ep, s->sl.data(), s->sl.size(), rw_write_cb, s);
      return;
    }
    s->write_done = 1;
  } else {
    s->write_err = error;
    s->write_done = 2;
  }
}

struct Count {
  int ok = 0, fail = 0;
  std::string last;
};
void count_cb(void* arg, const char* error) {
  Count* c = (Count*)arg;
  if (error) {
    c->fail++;
    c->last = error;
  } else {
    c->ok++;
  }
}

}  // namespace

extern "C" int drv_read_and_write(const b200_pair_ops* ops, uint64_t num_bytes, uint64_t write_size,
                                  uint64_t slice_size, int shutdown, int busy_us, int enable_poller,
                                  uint64_t* stats4) {
  Fixture f;
  if (!make_fixture(ops, busy_us, enable_poller, false, &f)) FAIL();
  RW s;
  s.read_ep = f.ep[0];
  s.write_ep = f.ep[1];
  s.target = num_bytes;
  s.cur_write = write_size;
  s.slice_size = slice_size;
  s.bytes_written = 0 - write_size;  // "pretend an initial write completed" (:266-270)
  rw_write_cb(&s, nullptr);
  b200_endpoint_read(s.read_ep, rw_read_cb, &s, 0);
  if (shutdown) {
    b200_endpoint_shutdown(s.read_ep, "Test Shutdown");
    b200_endpoint_shutdown(s.write_ep, "Test Shutdown");
  }
  const auto deadline = Clock::now() + std::chrono::seconds(120);
  while (!s.read_done || !s.write_done) {
    if (Clock::now() > deadline) FAIL();
    if (b200_engine_work(f.eng[0], 20) < 0) FAIL();
  }
  if (stats4) b200_engine_stats(f.eng[0], stats4);
  int rc = 0;
  if (s.bad) rc = __LINE__;
  if (!shutdown) {
    if (s.read_done != 1 || s.write_done != 1 || s.bytes_read != num_bytes) rc = __LINE__;
  } else {
    // the pending read must fail with the shutdown reason; the write either finished before the
    // shutdown took effect or failed with it
    if (s.read_done != 2 || s.read_err.find("Test Shutdown") == std::string::npos) rc = __LINE__;
    if (s.write_done == 2 && s.write_err.find("Test Shutdown") == std::string::npos) rc = __LINE__;
  }
  b200_endpoint_destroy(f.ep[0]);
  b200_endpoint_destroy(f.ep[1]);
  drop_fixture(&f);
  return rc;
}

extern "C" int drv_shutdown_sequence(const b200_pair_ops* ops, int busy_us) {
  Fixture f;
  if (!make_fixture(ops, busy_us, 0, false, &f)) FAIL();
  Count c;
  auto spin = [&](int n) {
    for (int i = 0; i < n; i++) b200_engine_work(f.eng[0], 1);
  };
  b200_endpoint_read(f.ep[0], count_cb, &c, 0);
  spin(3);
  if (c.fail != 0 || c.ok != 0) FAIL();
  b200_endpoint_shutdown(f.ep[0], "Test Shutdown");
  spin(3);
  if (c.fail != 1) FAIL();
  if (c.last.find("Test Shutdown") == std::string::npos || c.last.find("UNAVAILABLE") == std::string::npos) FAIL();
  b200_endpoint_read(f.ep[0], count_cb, &c, 0);  // a read after shutdown fails too
  spin(3);
  if (c.fail != 2) FAIL();
  // zero-length write on a shut-down endpoint: "EOF" (rdma_bp_posix.cc:566-574)
  b200_endpoint_write(f.ep[0], nullptr, 0, count_cb, &c);
  if (c.fail != 3 || c.last.find("EOF") == std::string::npos) FAIL();
  // a second shutdown is a no-op
  b200_endpoint_shutdown(f.ep[0], "Test Shutdown");
  spin(3);
  if (c.fail != 3) FAIL();
  // Reference behaviour kept: rdma_write does not consult the fd's shutdown state for a non-empty
  // buffer (rdma_bp_posix.cc:575-587), the pair still accepts it.
  const uint8_t a = 'a';
  b200_slice sl{&a, 1};
  b200_endpoint_write(f.ep[0], &sl, 1, count_cb, &c);
  spin(3);
  if (c.ok != 1 || c.fail != 3) FAIL();
  b200_endpoint_destroy(f.ep[0]);
  b200_endpoint_destroy(f.ep[1]);
  drop_fixture(&f);
  return 0;
}

extern "C" int drv_peer_close(const b200_pair_ops* ops, int busy_us, int enable_poller) {
  Fixture f;
  if (!make_fixture(ops, busy_us, enable_poller, false, &f)) FAIL();
  Count c;
  b200_endpoint_read(f.ep[0], count_cb, &c, 0);
  for (int i = 0; i < 3; i++) b200_engine_work(f.eng[0], 1);
  if (c.ok || c.fail) FAIL();
  b200_endpoint_destroy(f.ep[1]);  // rdma_free -> Disconnect -> peer_exit = 1 at the other end
  const auto deadline = Clock::now() + std::chrono::seconds(20);
  while (!c.fail && !c.ok) {
    if (Clock::now() > deadline) FAIL();
    b200_engine_work(f.eng[0], 5);
  }
  if (c.ok || c.fail != 1 || c.last.find("Pair closed") == std::string::npos) FAIL();
  // and a write of something that cannot complete reports the exit
  b200_endpoint_destroy(f.ep[0]);
  drop_fixture(&f);
  return 0;
}

// b200_engine_work must not hold the engine lock while it sleeps in epoll_wait (or runs callbacks):
// a thread inside b200_engine_work(timeout) and another one calling b200_endpoint_write / read on the same
// engine.  Every call of the second thread has to return long before the first one's wait ends.
extern "C" int drv_two_threads(const b200_pair_ops* ops, int wait_ms, int rounds, uint64_t* worst_call_us) {
  Fixture f;
  if (!make_fixture(ops, 0, 0, false, &f)) FAIL();
  std::atomic<int> stop{0}, in_work{0};
  std::thread worker([&] {
    while (!stop.load()) {
      in_work = 1;
      b200_engine_work(f.eng[0], wait_ms);  // busy window 0: goes straight to epoll_wait(wait_ms)
      in_work = 0;
    }
  });
  uint64_t worst = 0;
  int rc = 0;
  Count wc, rcnt;
  std::vector<uint8_t> block(4096, 0x42);
  b200_slice sl{block.data(), block.size()};
  for (int k = 0; k < rounds && !rc; k++) {
    while (!in_work.load()) {
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(2));  // the worker is asleep in epoll_wait now
    const auto t0 = Clock::now();
    b200_endpoint_write(f.ep[1], &sl, 1, count_cb, &wc);
    const auto t1 = Clock::now();
    b200_endpoint_read(f.ep[0], count_cb, &rcnt, 0);
    const auto t2 = Clock::now();
    const uint64_t us = std::chrono::duration_cast<std::chrono::microseconds>(std::max(t1 - t0, t2 - t1)).count();
    worst = std::max(worst, us);
    // both complete (the worker thread, or one of our own calls, runs the callbacks)
    const auto deadline = Clock::now() + std::chrono::seconds(20);
    while ((wc.ok + wc.fail < k + 1 || rcnt.ok + rcnt.fail < k + 1) && !rc) {
      if (Clock::now() > deadline) rc = __LINE__;
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
  }
  stop = 1;
  worker.join();
  if (!rc && (wc.fail || rcnt.fail)) rc = __LINE__;
  if (worst_call_us) *worst_call_us = worst;
  b200_endpoint_destroy(f.ep[0]);
  b200_endpoint_destroy(f.ep[1]);
  drop_fixture(&f);
  return rc;
}

namespace {

struct Stream {  // message framing on the byte stream: [u64 length][bytes]
  b200_endpoint* ep;
  b200_engine* eng;
  std::vector<uint8_t> inbuf;
  bool reading = false;
  std::string err;
  int write_busy = 0;
  std::vector<std::vector<uint8_t>> wblocks;
  std::vector<b200_slice> wsl;
};

void st_read_cb(void* arg, const char* error) {
  Stream* s = (Stream*)arg;
  s->reading = false;
  if (error) {
    s->err = error;
    return;
  }
  const b200_slice* in;
  size_t n = b200_endpoint_incoming(s->ep, &in);
  for (size_t i = 0; i < n; i++) s->inbuf.insert(s->inbuf.end(), (const uint8_t*)in[i].ptr, (const uint8_t*)in[i].ptr + in[i].len);
}
void st_write_cb(void* arg, const char* error) {
  Stream* s = (Stream*)arg;
  s->write_busy = 0;
  if (error) s->err = error;
}

// hand the message over the way chttp2 does: 9-byte frame headers + <= 16384-byte slices
void st_send(Stream* s, const std::vector<uint8_t>& msg) {
  s->wblocks.clear();
  s->wsl.clear();
  std::vector<uint8_t> hdr(8);
  uint64_t len = msg.size();
  memcpy(hdr.data(), &len, 8);
  s->wblocks.push_back(hdr);
  for (size_t off = 0; off < msg.size();) {
    size_t n = std::min<size_t>(16384, msg.size() - off);
    s->wblocks.push_back(std::vector<uint8_t>(9, 0x5a));
    s->wblocks.push_back(std::vector<uint8_t>(msg.begin() + off, msg.begin() + off + n));
    off += n;
  }
  for (auto& b : s->wblocks) s->wsl.push_back({b.data(), b.size()});
  s->write_busy = 1;
  b200_endpoint_write(s->ep, s->wsl.data(), s->wsl.size(), st_write_cb, s);
}

// returns false on error/timeout; strips the 9-byte pseudo frame headers again
bool st_recv(Stream* s, std::vector<uint8_t>* msg, Clock::time_point deadline) {
  auto need = [&](size_t n) {
    while (s->inbuf.size() < n || s->write_busy) {
      if (!s->err.empty() || Clock::now() > deadline) return false;
      if (s->inbuf.size() < n && !s->reading) {
        s->reading = true;
        b200_endpoint_read(s->ep, st_read_cb, s, 0);
      }
      if (s->inbuf.size() >= n && !s->write_busy) break;
      b200_engine_work(s->eng, 5);
    }
    return true;
  };
  if (!need(8)) return false;
  uint64_t len;
  memcpy(&len, s->inbuf.data(), 8);
  const uint64_t nfr = (len + 16383) / 16384;
  if (!need(8 + len + 9 * nfr)) return false;
  msg->resize(len);
  size_t pos = 8, out = 0;
  for (uint64_t fr = 0; fr < nfr; fr++) {
    pos += 9;
    size_t n = std::min<uint64_t>(16384, len - out);
    memcpy(msg->data() + out, s->inbuf.data() + pos, n);
    pos += n;
    out += n;
  }
  s->inbuf.erase(s->inbuf.begin(), s->inbuf.begin() + pos);
  return true;
}

uint64_t xorshift(uint64_t* st) {
  uint64_t x = *st;
  x ^= x >> 12;
  x ^= x << 25;
  x ^= x >> 27;
  *st = x;
  return x * 0x2545F4914F6CDD1DULL;
}

}  // namespace

extern "C" int drv_echo(const b200_pair_ops* ops, int n_msgs, uint64_t max_len, uint64_t seed, int busy_us,
                        int enable_poller, int threaded, uint64_t* bytes_out) {
  Fixture f;
  if (!make_fixture(ops, busy_us, enable_poller, threaded != 0, &f)) FAIL();
  Stream cli{f.ep[0], f.eng[0]}, srv{f.ep[1], f.eng[1]};
  std::atomic<int> srv_rc{0};
  const auto deadline = Clock::now() + std::chrono::seconds(300);
  auto serve_one = [&]() -> int {
    std::vector<uint8_t> m;
    if (!st_recv(&srv, &m, deadline)) return __LINE__;
    st_send(&srv, m);
    return 0;
  };
  std::thread server;
  if (threaded) {
    server = std::thread([&] {
      for (int i = 0; i < n_msgs; i++) {
        int rc = serve_one();
        if (rc) {
          srv_rc = rc;
          return;
        }
      }
      // flush the last reply
      while (srv.write_busy && srv.err.empty() && Clock::now() < deadline) b200_engine_work(srv.eng, 5);
    });
  }
  int rc = 0;
  uint64_t st = seed ? seed : 1, bytes = 0;
  for (int i = 0; i < n_msgs && !rc; i++) {
    const uint64_t len = 1 + xorshift(&st) % max_len;  // common.h:5-6: uniform in [1, max]
    std::vector<uint8_t> msg(len);
    for (uint64_t j = 0; j < len; j += 8) {
      uint64_t r = xorshift(&st);
      memcpy(msg.data() + j, &r, std::min<uint64_t>(8, len - j));
    }
    st_send(&cli, msg);
    std::vector<uint8_t> reply;
    if (threaded) {
      if (!st_recv(&cli, &reply, deadline)) rc = __LINE__;
    } else {
      // one thread plays both roles on one engine: the engine runs whichever callback is ready
      std::vector<uint8_t> m;
      if (!st_recv(&srv, &m, deadline)) rc = __LINE__;
      else {
        st_send(&srv, m);
        if (!st_recv(&cli, &reply, deadline)) rc = __LINE__;
      }
    }
    if (!rc && reply != msg) rc = __LINE__;  // GPR_ASSERT(msg == reply.message()), greeter_client.cc:61
    bytes += len;
  }
  if (threaded) server.join();
  if (!rc && srv_rc) rc = srv_rc;
  if (!rc && (!cli.err.empty() || !srv.err.empty())) rc = __LINE__;
  if (bytes_out) *bytes_out = bytes;
  b200_endpoint_destroy(f.ep[0]);
  b200_endpoint_destroy(f.ep[1]);
  drop_fixture(&f);
  return rc;
}

// Many connections on ONE engine pair (client engine, server engine): what a server pollset sees --
// `n_conns` fds in the busy-poll scan / epoll set, reads and writes of different connections
// interleaving in one pollable (ev_epollex_rdma_bpev_linux.cc:1104-1145 caps a pass at MAX_EPOLL_EVENTS
// = 100 synthesized events, so more than 100 ready connections need several passes).  Every client sends
// `rounds` messages of a connection-specific size and checks each echo.
extern "C" int drv_multi_echo(const b200_pair_ops* ops, int n_conns, int rounds, uint64_t max_len, uint64_t seed,
                              int busy_us, int enable_poller, int threaded, uint64_t* stats_out) {
  b200_engine* ce = b200_engine_create(ops, busy_us);
  b200_engine* se = threaded ? b200_engine_create(ops, busy_us) : ce;
  std::vector<b200_endpoint*> cep(n_conns), sep(n_conns);
  for (int c = 0; c < n_conns; c++) {
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) FAIL();
    std::thread t([&] { sep[c] = b200_endpoint_create(se, sv[1], "ipv4:server", enable_poller); });
    cep[c] = b200_endpoint_create(ce, sv[0], "ipv4:client", enable_poller);
    t.join();
    if (!cep[c] || !sep[c]) FAIL();
  }
  std::vector<Stream> cli, srv;
  for (int c = 0; c < n_conns; c++) {
    cli.push_back(Stream{cep[c], ce});
    srv.push_back(Stream{sep[c], se});
  }
  const auto deadline = Clock::now() + std::chrono::seconds(300);
  std::atomic<int> srv_rc{0};
  std::atomic<bool> stop{false};
  // server: poll all connections; whenever one has a complete message, echo it
  auto serve = [&](bool until_stop) -> int {
    std::vector<int> served(n_conns, 0);
    int total = 0;
    while (total < n_conns * rounds) {
      if (until_stop && stop.load()) break;
      if (Clock::now() > deadline) return __LINE__;
      bool progress = false;
      for (int c = 0; c < n_conns; c++) {
        Stream* s = &srv[c];
        if (!s->err.empty()) return __LINE__;
        if (s->write_busy) continue;
        if (s->inbuf.size() >= 8) {
          uint64_t len;
          memcpy(&len, s->inbuf.data(), 8);
          const uint64_t nfr = (len + 16383) / 16384;
          if (s->inbuf.size() >= 8 + len + 9 * nfr) {
            std::vector<uint8_t> m;
            if (!st_recv(s, &m, deadline)) return __LINE__;
            st_send(s, m);
            served[c]++;
            total++;
            progress = true;
            continue;
          }
        }
        if (!s->reading) {
          s->reading = true;
          b200_endpoint_read(s->ep, st_read_cb, s, 0);
        }
      }
      if (!progress) b200_engine_work(se, 2);
      if (!threaded) return 0;  // single-threaded mode: one sweep per call
    }
    while (threaded) {  // flush the last replies
      bool busy = false;
      for (auto& s : srv) busy = busy || s.write_busy;
      if (!busy || Clock::now() > deadline) break;
      b200_engine_work(se, 2);
    }
    return 0;
  };
  std::thread server;
  if (threaded) server = std::thread([&] { srv_rc = serve(false); });
  int rc = 0;
  uint64_t st = seed ? seed : 1;
  std::vector<std::vector<uint8_t>> sent(n_conns);
  for (int r = 0; r < rounds && !rc; r++) {
    for (int c = 0; c < n_conns; c++) {  // all connections have a request in flight at once
      const uint64_t len = 1 + (xorshift(&st) % max_len);
      sent[c].resize(len);
      for (uint64_t j = 0; j < len; j += 8) {
        uint64_t v = xorshift(&st);
        memcpy(sent[c].data() + j, &v, std::min<uint64_t>(8, len - j));
      }
      while (cli[c].write_busy && Clock::now() < deadline) b200_engine_work(ce, 1);
      st_send(&cli[c], sent[c]);
    }
    for (int c = 0; c < n_conns && !rc; c++) {
      std::vector<uint8_t> reply;
      if (threaded) {
        if (!st_recv(&cli[c], &reply, deadline)) rc = __LINE__;
      } else {
        // one thread: alternate server sweeps and client progress until this reply is complete
        while (!rc) {
          int s_rc = serve(false);
          if (s_rc) rc = s_rc;
          Stream* s = &cli[c];
          if (s->inbuf.size() >= 8) {
            uint64_t len;
            memcpy(&len, s->inbuf.data(), 8);
            if (s->inbuf.size() >= 8 + len + 9 * ((len + 16383) / 16384) && !s->write_busy) break;
          }
          if (!s->reading && s->err.empty()) {
            s->reading = true;
            b200_endpoint_read(s->ep, st_read_cb, s, 0);
          }
          b200_engine_work(ce, 1);
          if (Clock::now() > deadline || !s->err.empty()) rc = __LINE__;
        }
        if (!rc && !st_recv(&cli[c], &reply, deadline)) rc = __LINE__;
      }
      if (!rc && reply != sent[c]) rc = __LINE__;
    }
  }
  stop = true;
  if (threaded) server.join();
  if (!rc && srv_rc) rc = srv_rc;
  if (stats_out) b200_engine_stats(se, stats_out);
  for (int c = 0; c < n_conns; c++) {
    b200_endpoint_destroy(cep[c]);
    b200_endpoint_destroy(sep[c]);
  }
  if (se != ce) b200_engine_destroy(se);
  b200_engine_destroy(ce);
  return rc;
}
