/*
 * oracle_pair_ops.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A b200_pair_ops table (include/b200_endpoint.h) over the CPU oracle (oracle/rb_oracle.c), so the
 * `-m "not gpu"` tests can drive the product's endpoint state machine and BPEV poll loop
 * (grpc-rdma_b200/host/b200_endpoint.cc) without a GPU.  The product never links this file; the
 * GPU tests run the same driver with ops = NULL (the CUDA library).
 *
 * Single-threaded use only: the oracle's memcpy wire has no inter-thread ordering.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/eventfd.h>
#include <unistd.h>

#include "../../include/b200_endpoint.h"
#include "../../oracle/rb_oracle.h"

typedef struct {
  orb_pair* p;
  int efd;
  uint64_t ring;
} opair;

static uint64_t g_ring = 4096;
static int g_max_sge = 30;

void oracle_ops_config(uint64_t ring_bytes, int max_sge) {
  g_ring = ring_bytes;
  g_max_sge = max_sge;
}

static void* o_take(const char* id) {
  (void)id;
  opair* o = (opair*)calloc(1, sizeof(opair));
  o->efd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  o->ring = g_ring;
  return o;
}
static void o_putback(void* v) {
  opair* o = (opair*)v;
  if (o->p) orb_pair_destroy(o->p);
  close(o->efd);
  free(o);
}
static void o_init(void* v) {
  opair* o = (opair*)v;
  if (o->p) orb_pair_destroy(o->p);
  o->p = orb_pair_create(o->ring, g_max_sge);
}
static size_t o_addr(void* v, void* out48) {
  opair* o = (opair*)v;
  memset(out48, 0, B200_ADDRESS_BYTES);
  memcpy(out48, &o->p, sizeof(o->p));
  ((unsigned char*)out48)[32] = B200_PAIR_TAG_POLLABLE;
  memcpy((char*)out48 + 40, &o->ring, 8);
  return B200_ADDRESS_BYTES;
}
static int o_connect(void* v, const void* peer48, size_t n) {
  opair* o = (opair*)v;
  orb_pair* peer;
  uint64_t ring;
  if (n != B200_ADDRESS_BYTES) return 0;
  memcpy(&peer, peer48, sizeof(peer));
  memcpy(&ring, (const char*)peer48 + 40, 8);
  if (((const unsigned char*)peer48)[32] != B200_PAIR_TAG_POLLABLE || ring != o->ring) return 0; /* pair.cc:148-149 */
  o->p->peer = peer;
  o->p->status = ORB_CONNECTED;
  return 1;
}
static uint64_t o_send(void* v, const b200_slice* s, size_t n, size_t b) {
  return orb_pair_send(((opair*)v)->p, (const orb_slice*)s, n, b);
}
static uint64_t o_recv(void* v, void* d, uint64_t c) { return orb_pair_recv(((opair*)v)->p, d, c); }
static int o_has_msg(const void* v) { return orb_pair_has_message(((const opair*)v)->p); }
static int o_pending(const void* v) { return orb_pair_has_pending_writes(((const opair*)v)->p); }
static uint64_t o_readable(const void* v) { return orb_pair_readable(((const opair*)v)->p); }
static int o_status(void* v) { return orb_pair_get_status(((opair*)v)->p); }
static const char* o_error(const void* v) {
  (void)v;
  return "";
}
static int o_wfd(void* v) { return ((opair*)v)->efd; }
static void o_consume(void* v) {
  uint64_t x;
  if (read(((opair*)v)->efd, &x, 8) < 0) {
  }
}
static void o_disconnect(void* v) {
  opair* o = (opair*)v;
  if (o->p) orb_pair_disconnect(o->p);
}
static void o_noop(void* v) { (void)v; }

static const b200_pair_ops kOps = {o_take,   o_putback, o_init,    o_addr,     o_connect, o_send,
                                   o_recv,   o_has_msg, o_pending, o_readable, o_status,  o_error,
                                   o_wfd,    o_consume, o_disconnect, o_noop,  o_noop};

const b200_pair_ops* oracle_pair_ops(void) { return &kOps; }

/* ---- the B200-native widening of the table, over the same oracle: `submit` (one pass = the rdma_flush loops and
 * rdma_do_read loops of every queued endpoint, oracle's orb_pair_send_all / orb_pair_recv_drain) and the
 * completion-queue form (an op finishes at post time; poll just hands the count back).  With these the engine's
 * batching and pump code paths run on the CPU. */
static int o_submit(const b200_send_op* s, size_t ns, uint64_t* acc, const b200_recv_op* r, size_t nr, uint64_t* del,
                    int flags) {
  for (size_t i = 0; i < ns; i++) {
    orb_pair* p = ((opair*)s[i].pair)->p;
    acc[i] = (flags & B200_BATCH_UNTIL_BLOCKED) ? orb_pair_send_all(p, (const orb_slice*)s[i].slices, s[i].nslices, s[i].byte_idx, NULL)
                                                : orb_pair_send(p, (const orb_slice*)s[i].slices, s[i].nslices, s[i].byte_idx);
  }
  for (size_t i = 0; i < nr; i++) {
    orb_pair* p = ((opair*)r[i].pair)->p;
    del[i] = (flags & B200_BATCH_UNTIL_BLOCKED) ? orb_pair_recv_drain(p, r[i].dst, r[i].cap, NULL) : orb_pair_recv(p, r[i].dst, r[i].cap);
  }
  return 0;
}
static void* o_post_send(void* v, const b200_slice* s, size_t n, size_t b, int flags, int* again) {
  uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t));
  orb_pair* p = ((opair*)v)->p;
  *again = 0;
  *h = (flags & B200_BATCH_UNTIL_BLOCKED) ? orb_pair_send_all(p, (const orb_slice*)s, n, b, NULL) : orb_pair_send(p, (const orb_slice*)s, n, b);
  return h;
}
static void* o_post_recv(void* v, void* dst, uint64_t cap, int flags, int* again) {
  uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t));
  orb_pair* p = ((opair*)v)->p;
  *again = 0;
  *h = (flags & B200_BATCH_UNTIL_BLOCKED) ? orb_pair_recv_drain(p, dst, cap, NULL) : orb_pair_recv(p, dst, cap);
  return h;
}
static int o_poll(void* op, uint64_t* bytes) {
  *bytes = *(uint64_t*)op;
  free(op);
  return 1;
}
static const b200_pair_ops kOpsBatch = {o_take,   o_putback, o_init,    o_addr,     o_connect, o_send,
                                        o_recv,   o_has_msg, o_pending, o_readable, o_status,  o_error,
                                        o_wfd,    o_consume, o_disconnect, o_noop,  o_noop,    o_submit,
                                        NULL,     NULL,      NULL,      NULL,       NULL};
static const b200_pair_ops kOpsAsync = {o_take,   o_putback, o_init,    o_addr,     o_connect, o_send,
                                        o_recv,   o_has_msg, o_pending, o_readable, o_status,  o_error,
                                        o_wfd,    o_consume, o_disconnect, o_noop,  o_noop,    o_submit,
                                        NULL,     NULL,      o_post_send, o_post_recv, o_poll};
const b200_pair_ops* oracle_pair_ops_batch(void) { return &kOpsBatch; }
const b200_pair_ops* oracle_pair_ops_async(void) { return &kOpsAsync; }
