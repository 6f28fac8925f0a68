/*
 * rb_oracle.h -- CPU restatement (plain C) of the reference's RDMA_BPEV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under grpc-rdma_b200/ may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg use it, and only as the checker / the
 * timed CPU baseline -- never as the product path.
 *
 * Parity pin: this port is validated against the reference's own code
 * (ring_buffer.cc + pair.cc compiled UNMODIFIED over a loopback fake
 * libibverbs into oracle/_ref/libref_pair.so, see oracle/Makefile and
 * oracle/ref_harness.cc) and against tests/golden/ vectors generated from that
 * build (tests/golden/make_golden.py).  The reference's own unit tests for this
 * path (test/core/ibverbs/) are absent from the snapshot, so the compiled
 * reference is the pin.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference root, src/core/lib/ibverbs/ unless noted).
 */
#ifndef RB_ORACLE_H
#define RB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_ALIGN 8u           /* ring_buffer.h:49  alignment = sizeof(tag_t)       */
#define ORB_RESERVED 24u       /* ring_buffer.h:52  reserved_space = 3*alignment    */
#define ORB_FOOTER UINT64_MAX  /* ring_buffer.h:50  footer = all ones               */

/* pair.h:44-51 PairStatus */
enum orb_status {
  ORB_UNINITIALIZED = 0,
  ORB_INITIALIZED = 1,
  ORB_CONNECTED = 2,
  ORB_HALF_CLOSED = 3,
  ORB_DISCONNECTED = 4,
  ORB_ERROR = 5
};

typedef struct {
  const uint8_t* ptr;
  uint64_t len;
} orb_slice; /* flattened grpc_slice: GRPC_SLICE_START_PTR/LENGTH, include/grpc/slice.h */

/* ring_buffer.h:203-208 */
typedef struct {
  uint8_t* buf;
  uint64_t capacity;
  uint64_t mask;
  uint64_t head;
  uint64_t moving_head;
  uint64_t remain;
} orb_ring;

/* pair.h:100-103 status_report */
typedef struct {
  uint64_t remote_head;
  int32_t peer_exit;
  int32_t _pad;
} orb_status_report;

typedef struct orb_pair {
  orb_ring ring;               /* recv_buffers_[kDataBuffer] + ring_buf_      */
  uint8_t* staging;            /* send_buffers_[kDataBuffer], capacity/2      */
  uint64_t staging_size;
  orb_status_report status_in; /* recv_buffers_[kStatusBuffer]: written by peer */
  orb_status_report status_out;/* send_buffers_[kStatusBuffer]                */
  uint64_t remote_tail;        /* pair.h remote_tail_                         */
  uint64_t internal_read_size; /* pair.h internal_read_size_                  */
  int partial_write;           /* pair.h partial_write_                       */
  int status;                  /* enum orb_status                             */
  int max_sge;                 /* ibv_device_attr.max_sge (30 on the authors' mlx5) */
  struct orb_pair* peer;       /* the wire: memcpy into peer->ring.buf         */
  uint64_t total_read, total_write;
  uint64_t n_status_writes;    /* number of credit (status) writes posted      */
} orb_pair;

/* ---- integer helpers (ring_buffer.h:180-189, 232-248) ---- */
uint64_t orb_round_up(uint64_t v);
uint64_t orb_round_down(uint64_t v);
uint64_t orb_encoded_size(uint64_t payload);            /* GetEncodedSize        */
uint64_t orb_calc_writable(uint64_t space);             /* CalculateWritableSize */
uint64_t orb_free_size(uint64_t cap, uint64_t head, uint64_t tail);     /* ring_buffer.cc:99  */
uint64_t orb_writable_size(uint64_t cap, uint64_t head, uint64_t tail); /* ring_buffer.cc:106 */

/* ---- ring (receiver side) ---- */
void orb_ring_init(orb_ring* r, uint8_t* buf, uint64_t capacity); /* ring_buffer.cc:21-26,50-55 */
int orb_ring_has_message(const orb_ring* r);                      /* ring_buffer.cc:56-65  */
uint64_t orb_ring_readable(const orb_ring* r);                    /* ring_buffer.cc:67-97  */
uint64_t orb_ring_read(orb_ring* r, void* dst, uint64_t cap,
                       uint64_t* internal_bytes_read);            /* ring_buffer.cc:122-191 */
/* writer-side frame placement: copy `len` encoded bytes to ring at `tail`
 * with wrap (what the RDMA WRs of GetWriteRequests do, ring_buffer.cc:261-330);
 * returns new tail. */
uint64_t orb_ring_place(uint8_t* ring_buf, uint64_t capacity, uint64_t tail,
                        const uint8_t* encoded, uint64_t len);

/* ---- pair ---- */
orb_pair* orb_pair_create(uint64_t ring_capacity, int max_sge);  /* pair.cc:85-141 Init */
void orb_pair_destroy(orb_pair* p);
void orb_pair_connect(orb_pair* a, orb_pair* b);                 /* pair.cc:143-168     */
uint64_t orb_pair_send(orb_pair* p, const orb_slice* slices, size_t n,
                       size_t byte_idx);                         /* pair.cc:645-734     */
uint64_t orb_pair_recv(orb_pair* p, void* dst, uint64_t cap);    /* pair.cc:264-286     */
int orb_pair_has_message(const orb_pair* p);                     /* pair.cc:288         */
int orb_pair_has_pending_writes(const orb_pair* p);              /* pair.cc:303         */
uint64_t orb_pair_readable(const orb_pair* p);                   /* pair.cc:290-292     */
uint64_t orb_pair_writable(const orb_pair* p);                   /* pair.cc:294-301     */
int orb_pair_get_status(const orb_pair* p);                      /* pair.cc:349-375     */
void orb_pair_disconnect(orb_pair* p);                           /* pair.cc:325-347     */

/* Endpoint-level helpers (rdma_bp_posix.cc): drive Send until it accepts
 * nothing more (rdma_flush :470-524 re-entered from rdma_handle_write :527),
 * and Recv until the destination is full or the ring has no complete frame
 * (rdma_do_read :180-286).  Return payload bytes moved; *calls = number of
 * Send/Recv invocations that moved >0 bytes. */
uint64_t orb_pair_send_all(orb_pair* p, const orb_slice* slices, size_t n,
                           size_t byte_idx, uint64_t* calls);
uint64_t orb_pair_recv_drain(orb_pair* p, void* dst, uint64_t cap, uint64_t* calls);

/* Multi-threaded CPU baseline over this port: `conns` loopback connections,
 * each sending `warm` untimed then `msgs` timed messages built from the chttp2-shaped slice list
 * (lens[0..nslices)) out of `src` and draining them into `dst`; `threads`
 * pthreads each own conns/threads connections.  Returns elapsed seconds
 * (wall), fills *delivered with payload bytes delivered. */
double orb_bench_stream(int conns, int threads, int warm, int msgs, uint64_t ring_capacity,
                        const uint64_t* lens, size_t nslices, uint64_t* delivered,
                        uint64_t* checksum);

#ifdef __cplusplus
}
#endif
#endif
