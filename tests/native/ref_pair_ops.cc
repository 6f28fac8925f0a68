// ref_pair_ops.cc -- TEST INFRASTRUCTURE ONLY.
//
// A b200_pair_ops table (include/b200_endpoint.h) over the REFERENCE's own PairPollable / Poller
// (oracle/_ref/libref_pair_dbg.so: pair.cc, ring_buffer.cc, poller.cc compiled unmodified from
// /root/reference with asserts ON, over the loopback fake verbs).  It lets the CPU tests drive the
// product's endpoint state machine and BPEV poll loop (grpc-rdma_b200/host/b200_endpoint.cc) against the
// real reference pair: the reference's invariants (ContentAssertion, ring_buffer.cc asserts) police
// every Send / Recv the endpoint issues, from two threads, with the reference's Poller kicking eventfds.
#include <stdint.h>
#include <string.h>
#include <sys/eventfd.h>
#include <unistd.h>

#include "../../include/b200_endpoint.h"

extern "C" {
void ref_set_ring_kb(uint32_t kb);
void* ref_pair_create();
void ref_pair_destroy(void* p);
int ref_pair_connect_to(void* p, const void* addr, size_t n);
size_t ref_pair_address(void* p, void* out, size_t cap);
uint64_t ref_pair_send(void* p, const b200_slice* slices, size_t n, size_t byte_idx);
uint64_t ref_pair_recv(void* p, void* dst, uint64_t cap);
int ref_pair_has_message(void* p);
int ref_pair_has_pending_writes(void* p);
uint64_t ref_pair_readable(void* p);
int ref_pair_get_status(void* p);
const char* ref_pair_error(void* p);
void ref_pair_disconnect(void* p);
int ref_pair_wakeup_fd(void* p);
void ref_poller_add(void* p);
void ref_poller_remove(void* p);
}

namespace {
void* r_take(const char*) { return ref_pair_create(); }  // new PairPollable + Init (pair.cc:85)
void r_putback(void* p) { ref_pair_destroy(p); }
void r_init(void*) {}  // Init() ran in take; a second Init from INITIALIZED is a no-op in the reference too
size_t r_addr(void* p, void* out) { return ref_pair_address(p, out, B200_ADDRESS_BYTES); }
int r_connect(void* p, const void* a, size_t n) { return ref_pair_connect_to(p, a, n); }
uint64_t r_send(void* p, const b200_slice* s, size_t n, size_t b) { return ref_pair_send(p, s, n, b); }
uint64_t r_recv(void* p, void* d, uint64_t c) { return ref_pair_recv(p, d, c); }
int r_has_msg(const void* p) { return ref_pair_has_message(const_cast<void*>(p)); }
int r_pending(const void* p) { return ref_pair_has_pending_writes(const_cast<void*>(p)); }
uint64_t r_readable(const void* p) { return ref_pair_readable(const_cast<void*>(p)); }
int r_status(void* p) { return ref_pair_get_status(p); }
const char* r_error(const void* p) { return ref_pair_error(const_cast<void*>(p)); }
int r_wfd(void* p) { return ref_pair_wakeup_fd(p); }
void r_consume(void* p) {
  uint64_t v;
  if (read(ref_pair_wakeup_fd(p), &v, 8) < 0) {
  }
}
void r_disconnect(void* p) { ref_pair_disconnect(p); }
void r_padd(void* p) { ref_poller_add(p); }
void r_premove(void* p) { ref_poller_remove(p); }
const b200_pair_ops kOps = {r_take, r_putback, r_init,    r_addr,     r_connect, r_send,  r_recv, r_has_msg, r_pending,
                            r_readable, r_status, r_error, r_wfd, r_consume, r_disconnect, r_padd, r_premove};
}  // namespace

extern "C" const b200_pair_ops* ref_pair_ops(void) { return &kOps; }
extern "C" void ref_ops_config(uint32_t ring_kb) { ref_set_ring_kb(ring_kb); }
