/*
 * rb_oracle.c -- CPU restatement of the reference RDMA_BPEV ring/pair hot path.
 * TEST INFRASTRUCTURE ONLY (see rb_oracle.h).  Parity pin: oracle/_ref (the
 * reference's own ring_buffer.cc + pair.cc built over a loopback fake verbs)
 * and tests/golden/.
 */
#define _GNU_SOURCE
#include "rb_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ helpers */

/* ring_buffer.h:232-239 round_up */
uint64_t orb_round_up(uint64_t v) { return (v + (ORB_ALIGN - 1)) & ~(uint64_t)(ORB_ALIGN - 1); }
/* ring_buffer.h:242-248 round_down */
uint64_t orb_round_down(uint64_t v) { return v & ~(uint64_t)(ORB_ALIGN - 1); }
/* ring_buffer.h:180-183 GetEncodedSize: header + padded payload + footer */
uint64_t orb_encoded_size(uint64_t payload) { return 2u * ORB_ALIGN + orb_round_up(payload); }
/* ring_buffer.h:185-189 CalculateWritableSize */
uint64_t orb_calc_writable(uint64_t space) {
  if (space <= ORB_RESERVED) return 0;
  return orb_round_down(space - ORB_RESERVED);
}
/* ring_buffer.cc:99-104 GetFreeSize */
uint64_t orb_free_size(uint64_t cap, uint64_t head, uint64_t tail) {
  uint64_t used = (tail + cap - head) & (cap - 1);
  return cap - used;
}
/* ring_buffer.cc:106-116 GetWritableSize(head, tail) */
uint64_t orb_writable_size(uint64_t cap, uint64_t head, uint64_t tail) {
  uint64_t f = orb_free_size(cap, head, tail);
  return f > ORB_RESERVED ? f - ORB_RESERVED : 0;
}

static uint64_t ld64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}
static void st64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

/* --------------------------------------------------------------------- ring */

/* ring_buffer.cc:21-26 (ctor: power-of-two capacity) + :50-55 Init (zero fill) */
void orb_ring_init(orb_ring* r, uint8_t* buf, uint64_t capacity) {
  r->buf = buf;
  r->capacity = capacity;
  r->mask = capacity - 1;
  r->head = r->moving_head = r->remain = 0;
  memset(buf, 0, capacity);
}

/* ring_buffer.cc:56-65 */
int orb_ring_has_message(const orb_ring* r) {
  if (r->remain > 0) return 1;
  return ld64(r->buf + r->head) != 0;
}

/* ring_buffer.cc:67-97.  The reference spins (goto retry) while the header
 * reads as > capacity-24 (torn read); a single-threaded oracle reports such a
 * header as "nothing readable". */
uint64_t orb_ring_readable(const orb_ring* r) {
  if (r->remain > 0) return r->remain;
  uint64_t size = ld64(r->buf + r->head);
  if (size == 0) return 0;
  if (size > r->capacity - ORB_RESERVED) return 0;
  uint64_t foot = (r->head + ORB_ALIGN + orb_round_up(size)) & r->mask;
  return ld64(r->buf + foot) == ORB_FOOTER ? size : 0;
}

/* ring_buffer.cc:122-191 Read: deliver up to `cap` bytes of the frame at head
 * (or the rest of a partially consumed frame) and zero everything retired. */
uint64_t orb_ring_read(orb_ring* r, void* dst, uint64_t cap, uint64_t* internal) {
  uint64_t readable = orb_ring_readable(r);
  uint64_t n = readable < cap ? readable : cap;
  uint64_t prev_mh = r->moving_head;
  if (n == 0) {
    if (internal) *internal = 0;
    return 0;
  }
  if (r->remain == 0) { /* first touch of this frame (:135-147) */
    st64(r->buf + r->head, 0);
    r->moving_head = (r->head + ORB_ALIGN) & r->mask;
    r->head = (r->head + 2u * ORB_ALIGN + orb_round_up(readable)) & r->mask;
  }
  uint64_t end = (r->moving_head + n) & r->mask;
  uint64_t seg1 = n, seg2 = 0;
  if (!(r->moving_head < end)) { /* wraps (:152-157) */
    seg2 = end;
    seg1 = n - seg2;
  }
  memcpy(dst, r->buf + r->moving_head, seg1);
  memset(r->buf + r->moving_head, 0, seg1);
  if (seg2) {
    memcpy((uint8_t*)dst + seg1, r->buf, seg2);
    memset(r->buf, 0, seg2);
  }
  r->moving_head = (r->moving_head + n) & r->mask;
  r->remain = readable - n;
  if (r->remain == 0) { /* frame finished: pad + footer (:170-183) */
    uint64_t up = orb_round_up(r->moving_head);
    for (uint64_t pos = r->moving_head; pos < up; pos++) r->buf[pos & r->mask] = 0;
    r->moving_head = up & r->mask;
    st64(r->buf + r->moving_head, 0);
    r->moving_head = (r->moving_head + ORB_ALIGN) & r->mask;
  }
  if (internal) *internal = (r->moving_head + r->capacity - prev_mh) & r->mask;
  return n;
}

/* What the 1-2 RDMA WRs built by GetWriteRequests (ring_buffer.cc:261-330)
 * do to the remote ring: a byte run starting at `tail`, split at the ring end. */
uint64_t orb_ring_place(uint8_t* ring_buf, uint64_t capacity, uint64_t tail,
                        const uint8_t* encoded, uint64_t len) {
  uint64_t first = capacity - tail;
  if (first > len) first = len;
  memcpy(ring_buf + tail, encoded, first);
  if (len > first) memcpy(ring_buf, encoded + first, len - first);
  return (tail + len) & (capacity - 1);
}

/* --------------------------------------------------------------------- pair */

/* pair.cc:85-141 Init: ring = cap, staging = cap/2, status bufs zeroed */
orb_pair* orb_pair_create(uint64_t cap, int max_sge) {
  if (cap <= ORB_RESERVED || (cap & (cap - 1)) != 0) return NULL; /* ring_buffer.cc:22-23 */
  orb_pair* p = (orb_pair*)calloc(1, sizeof(orb_pair));
  uint8_t* ring = (uint8_t*)malloc(cap);
  p->staging_size = cap / 2;
  p->staging = (uint8_t*)calloc(1, p->staging_size);
  orb_ring_init(&p->ring, ring, cap);
  p->max_sge = max_sge;
  p->status = ORB_INITIALIZED;
  return p;
}

void orb_pair_destroy(orb_pair* p) {
  if (!p) return;
  free(p->ring.buf);
  free(p->staging);
  free(p);
}

/* pair.cc:143-168 Connect (tag / ring-size equality asserted there) */
void orb_pair_connect(orb_pair* a, orb_pair* b) {
  a->peer = b;
  b->peer = a;
  a->status = b->status = ORB_CONNECTED;
}

/* pair.cc:349-375 get_status (QP liveness probe has no loopback analogue) */
int orb_pair_get_status(const orb_pair* p) {
  if (p->status == ORB_CONNECTED && p->status_in.peer_exit == 1) return ORB_HALF_CLOSED;
  return p->status;
}

/* pair.cc:645-734 Send(grpc_slice*, n, byte_idx): one frame per slice, a slice
 * is cut only when staging or remote credit runs out; at most max_sge frames. */
uint64_t orb_pair_send(orb_pair* p, const orb_slice* slices, size_t n, size_t byte_idx) {
  if (p->status != ORB_CONNECTED) return 0; /* :657 (status_, not get_status()) */
  uint64_t cap = p->ring.capacity;
  uint64_t rh = p->status_in.remote_head; /* credit snapshot, once (:650) */
  uint64_t rt = p->remote_tail;
  uint64_t st = 0, total = 0, written = 0;
  int nsge = 0;
  for (size_t i = 0; i < n; i++) total += slices[i].len;
  total -= byte_idx;
  for (size_t i = 0; i < n && nsge < p->max_sge; i++) {
    const uint8_t* ptr = slices[i].ptr + byte_idx;
    uint64_t len = slices[i].len - byte_idx;
    byte_idx = 0;
    uint64_t a = orb_calc_writable(p->staging_size - st);
    uint64_t b = orb_calc_writable(orb_free_size(cap, rh, rt));
    uint64_t pay = len;
    if (a < pay) pay = a;
    if (b < pay) pay = b;
    if (pay == 0) break; /* :683-685 (also stops on a zero-length slice) */
    uint8_t* f = p->staging + st; /* frame: [len][payload][pad as-is][~0] */
    st64(f, pay);
    memcpy(f + ORB_ALIGN, ptr, pay);
    st64(f + ORB_ALIGN + orb_round_up(pay), ORB_FOOTER);
    uint64_t e = orb_encoded_size(pay);
    st += e;
    rt = (rt + e) & (cap - 1);
    written += pay;
    nsge++;
  }
  p->partial_write = written < total; /* :712 */
  if (nsge > 0) {                     /* the wire: RDMA write(s) into the peer ring */
    p->remote_tail = orb_ring_place(p->peer->ring.buf, cap, p->remote_tail, p->staging, st);
  }
  p->total_write += written;
  return written;
}

/* pair.cc:264-286 Recv + :624-641 updateStatus: return credit when >= cap/2 retired */
uint64_t orb_pair_recv(orb_pair* p, void* dst, uint64_t cap) {
  if (p->status != ORB_CONNECTED) return 0;
  uint64_t internal = 0;
  uint64_t n = orb_ring_read(&p->ring, dst, cap, &internal);
  p->internal_read_size += internal;
  p->total_read += n;
  if (p->internal_read_size >= p->ring.capacity / 2) {
    p->status_out.remote_head = p->ring.moving_head;
    p->peer->status_in = p->status_out; /* 16-byte RDMA write of status_report */
    p->n_status_writes++;
    p->internal_read_size = 0;
  }
  return n;
}

int orb_pair_has_message(const orb_pair* p) { return orb_ring_has_message(&p->ring); }
int orb_pair_has_pending_writes(const orb_pair* p) { return p->partial_write; }
uint64_t orb_pair_readable(const orb_pair* p) {
  return p->status == ORB_CONNECTED ? orb_ring_readable(&p->ring) : 0;
}
uint64_t orb_pair_writable(const orb_pair* p) {
  return orb_writable_size(p->ring.capacity, p->status_in.remote_head, p->remote_tail);
}

/* pair.cc:325-347 Disconnect: tell the peer we are leaving */
void orb_pair_disconnect(orb_pair* p) {
  if (p->status == ORB_UNINITIALIZED || p->status == ORB_DISCONNECTED) return;
  if (orb_pair_get_status(p) == ORB_CONNECTED && p->peer) {
    p->status_out.peer_exit = 1;
    p->status_out.remote_head = p->ring.moving_head; /* updateStatus :628 */
    p->peer->status_in = p->status_out;
    p->n_status_writes++;
  }
  p->status = ORB_DISCONNECTED;
}

/* rdma_bp_posix.cc:470-524 rdma_flush, re-entered via rdma_handle_write :527
 * while the pair keeps accepting bytes. */
uint64_t orb_pair_send_all(orb_pair* p, const orb_slice* slices, size_t n, size_t byte_idx,
                           uint64_t* calls) {
  uint64_t sent_total = 0, ncalls = 0;
  size_t idx = 0;
  while (idx < n) {
    uint64_t sent = orb_pair_send(p, slices + idx, n - idx, byte_idx);
    if (sent == 0) break;
    ncalls++;
    sent_total += sent;
    while (sent > 0) { /* :480-493 advance slice / byte cursor */
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
  return sent_total;
}

/* rdma_bp_posix.cc:180-286 rdma_do_read: Recv repeatedly into what is left of dst */
uint64_t orb_pair_recv_drain(orb_pair* p, void* dst, uint64_t cap, uint64_t* calls) {
  uint64_t got = 0, ncalls = 0;
  while (got < cap) {
    uint64_t n = orb_pair_recv(p, (uint8_t*)dst + got, cap - got);
    if (n == 0) break;
    got += n;
    ncalls++;
  }
  if (calls) *calls = ncalls;
  return got;
}

/* ------------------------------------------------- multi-threaded CPU baseline */

typedef struct {
  int first_conn, n_conn, warm, msgs;
  uint64_t ring_capacity;
  const uint64_t* lens;
  size_t nslices;
  uint64_t delivered, checksum;
  pthread_barrier_t* start;
} orb_worker;

static uint64_t fnv1a(const uint8_t* p, uint64_t n, uint64_t h) {
  for (uint64_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}

static void* orb_worker_main(void* arg) {
  orb_worker* w = (orb_worker*)arg;
  uint64_t msg_bytes = 0;
  for (size_t i = 0; i < w->nslices; i++) msg_bytes += w->lens[i];
  int nc = w->n_conn;
  orb_pair** tx = (orb_pair**)calloc(nc, sizeof(*tx));
  orb_pair** rx = (orb_pair**)calloc(nc, sizeof(*rx));
  uint8_t** src = (uint8_t**)calloc(nc, sizeof(*src));
  uint8_t** dst = (uint8_t**)calloc(nc, sizeof(*dst));
  orb_slice* sl = (orb_slice*)calloc(w->nslices, sizeof(*sl));
  for (int c = 0; c < nc; c++) {
    tx[c] = orb_pair_create(w->ring_capacity, 30);
    rx[c] = orb_pair_create(w->ring_capacity, 30);
    orb_pair_connect(tx[c], rx[c]);
    src[c] = (uint8_t*)malloc(msg_bytes);
    dst[c] = (uint8_t*)malloc(msg_bytes);
    for (uint64_t i = 0; i < msg_bytes; i++)
      src[c][i] = (uint8_t)(i + 131u * (unsigned)(w->first_conn + c));
    memset(dst[c], 0, msg_bytes);
  }
  uint64_t delivered = 0;
  for (int m = -w->warm; m < w->msgs; m++) {
    if (m == 0) {
      pthread_barrier_wait(w->start); /* warm-up messages done: timing starts */
      delivered = 0;
    }
    for (int c = 0; c < nc; c++) {
      uint64_t off = 0;
      for (size_t i = 0; i < w->nslices; i++) {
        sl[i].ptr = src[c] + off;
        sl[i].len = w->lens[i];
        off += w->lens[i];
      }
      /* closed loop: flush as far as credit allows, drain, repeat */
      size_t idx = 0, bidx = 0;
      uint64_t got = 0;
      while (got < msg_bytes) {
        while (idx < w->nslices) {
          uint64_t sent = orb_pair_send(tx[c], sl + idx, w->nslices - idx, bidx);
          if (sent == 0) break;
          while (sent > 0) {
            uint64_t left = sl[idx].len - bidx;
            if (sent >= left) { sent -= left; idx++; bidx = 0; }
            else { bidx += sent; sent = 0; }
          }
        }
        got += orb_pair_recv_drain(rx[c], dst[c] + got, msg_bytes - got, NULL);
      }
      delivered += got;
    }
  }
  pthread_barrier_wait(w->start);
  uint64_t h = 0xcbf29ce484222325ULL;
  for (int c = 0; c < nc; c++) h = fnv1a(dst[c], msg_bytes, h);
  w->delivered = delivered;
  w->checksum = h;
  for (int c = 0; c < nc; c++) {
    orb_pair_destroy(tx[c]);
    orb_pair_destroy(rx[c]);
    free(src[c]);
    free(dst[c]);
  }
  free(tx); free(rx); free(src); free(dst); free(sl);
  return NULL;
}

double orb_bench_stream(int conns, int threads, int warm, int msgs, uint64_t ring_capacity,
                        const uint64_t* lens, size_t nslices, uint64_t* delivered,
                        uint64_t* checksum) {
  if (threads < 1) threads = 1;
  if (threads > conns) threads = conns;
  pthread_t* th = (pthread_t*)calloc(threads, sizeof(*th));
  orb_worker* ws = (orb_worker*)calloc(threads, sizeof(*ws));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, threads + 1);
  int base = 0;
  for (int t = 0; t < threads; t++) {
    int nc = conns / threads + (t < conns % threads ? 1 : 0);
    ws[t].first_conn = base;
    ws[t].n_conn = nc;
    ws[t].warm = warm;
    ws[t].msgs = msgs;
    ws[t].ring_capacity = ring_capacity;
    ws[t].lens = lens;
    ws[t].nslices = nslices;
    ws[t].start = &bar;
    base += nc;
    pthread_create(&th[t], NULL, orb_worker_main, &ws[t]);
  }
  struct timespec t0, t1;
  pthread_barrier_wait(&bar); /* all workers allocated + initialised */
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&bar); /* all workers done streaming */
  clock_gettime(CLOCK_MONOTONIC, &t1);
  uint64_t d = 0, h = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    d += ws[t].delivered;
    h ^= ws[t].checksum;
  }
  pthread_barrier_destroy(&bar);
  free(th);
  free(ws);
  if (delivered) *delivered = d;
  if (checksum) *checksum = h;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
