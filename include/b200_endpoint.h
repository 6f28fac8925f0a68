/*
 * b200_endpoint.h -- host side of the RDMA_BPEV path above the pair ABI: the endpoint state
 * machine and the busy-poll/epoll hybrid completion loop, as a C ABI.
 *
 * Mirrors, for the reference (pwrliang/grpc-rdma, paths relative to its root):
 *   grpc_rdma endpoint        src/core/lib/iomgr/rdma_bp_posix.cc
 *     rdma_read :343  rdma_handle_read :328  rdma_continue_read :306  rdma_do_read :180
 *     rdma_write :559 rdma_flush :470        rdma_handle_write :527
 *     rdma_shutdown :100  rdma_destroy :168  exchange_data :640  grpc_rdma_bp_create :706
 *   BPEV engine poll loop     src/core/lib/iomgr/ev_epollex_rdma_bpev_linux.cc
 *     pollable_epoll :1079 (busy-poll window :1104-1145, epoll_wait :1153-1163)
 *     pollable_process_events :977 (eventfd tag events :1010-1035, synthetic events :1036-1066)
 *     pollable_add_fd :708-749 (eventfd registered EPOLLIN|EPOLLET)  fd_orphan :524-548
 *
 * It is the part of grpc_endpoint_vtable (endpoint.h:42-57) that has behaviour: one outstanding
 * read and one outstanding write, edge-triggered re-arm, `inq` hint, the slice the endpoint
 * allocates itself (max(256, readable)), trimming into last_read_buffer, zero-length writes
 * completing inline, and the error strings ("Pair closed", "Pair error, ...", "Peer has been
 * exited", "EOF").  gRPC's closures/slice buffers/resource quota are replaced by a callback and
 * flat b200_slice arrays; nothing of gRPC is linked.
 *
 * Implementation: grpc-rdma_b200/host/b200_endpoint.cc -> lib/libb200_endpoint.so.  The pair
 * operations come from a vtable; NULL selects the CUDA library (libb200rdma.so).  Tests may pass
 * another table to drive the same state machine over the CPU oracle (host-logic tests).
 */
#ifndef B200_ENDPOINT_H
#define B200_ENDPOINT_H

#include <stddef.h>
#include <stdint.h>

#include "b200_pair.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_engine b200_engine;     /* one pollable: epoll set + the rdma fd list */
typedef struct b200_endpoint b200_endpoint; /* grpc_rdma, rdma_bp_posix.cc:45-83 */

/* grpc_closure stand-in: error == NULL means GRPC_ERROR_NONE. */
typedef void (*b200_closure_fn)(void* arg, const char* error);

/* What the endpoint and the engine call on a pair (the side-face of SURVEY.md section 8b). */
typedef struct b200_pair_ops {
  void* (*pool_take)(const char* id);
  void (*pool_putback)(void* pair);
  void (*init)(void* pair);
  size_t (*self_address)(void* pair, void* out48);
  int (*connect)(void* pair, const void* peer48, size_t n);
  uint64_t (*send)(void* pair, const b200_slice* slices, size_t n, size_t byte_idx);
  uint64_t (*recv)(void* pair, void* dst, uint64_t cap);
  int (*has_message)(const void* pair);
  int (*has_pending_writes)(const void* pair);
  uint64_t (*readable)(const void* pair);
  int (*status)(void* pair); /* enum b200_status */
  const char* (*error)(const void* pair);
  int (*wakeup_read_fd)(void* pair);
  void (*consume_wakeup)(void* pair);
  void (*disconnect)(void* pair);
  void (*poller_add)(void* pair);
  void (*poller_remove)(void* pair);
  /* ---- optional (NULL = not provided): the B200-native widening.
   * submit: one event-loop pass -- every ready Send (rdma_flush loop) and Recv (rdma_do_read loop) posted
   * together and waited for together (b200_pairs_submit).  When present the engine BATCHES: reads and writes
   * are queued by rdma_read / rdma_write / the readiness scan and executed by the next b200_engine_work pass.
   * mem_alloc / mem_free: GPU-addressable memory for the read slices the endpoint allocates itself
   * (rdma_bp_posix.cc:308-317); they are pooled and 256-byte aligned. */
  int (*submit)(const b200_send_op* sops, size_t ns, uint64_t* accepted, const b200_recv_op* rops, size_t nr,
                uint64_t* delivered, int flags);
  void* (*mem_alloc)(size_t bytes);
  void (*mem_free)(void* p);
  /* completion-queue form (b200_pair_post_send / post_recv / b200_async_poll): when present the engine never
   * waits for the GPU -- each pass polls what is in flight and posts what is queued.  post returns NULL with
   * *again = 1 (no free queue entry: retry next pass), 2 (not available right now: run call by call), 0 (error). */
  void* (*post_send)(void* pair, const b200_slice* slices, size_t n, size_t byte_idx, int flags, int* again);
  void* (*post_recv)(void* pair, void* dst, uint64_t cap, int flags, int* again);
  int (*poll)(void* op, uint64_t* bytes);
} b200_pair_ops;

/* NULL ops = libb200rdma.so.  busy_poll_us < 0 = GRPC_RDMA_BUSY_POLLING_TIMEOUT_US (500). */
b200_engine* b200_engine_create(const b200_pair_ops* ops, int busy_poll_us);
void b200_engine_destroy(b200_engine* e);
/* pollset_work: one pass of pollable_epoll + pollable_process_events.  Scans the pairs for up
 * to the busy-poll window, then falls back to epoll_wait(timeout_ms) on their eventfds.
 * Returns the number of events handled (callbacks run inside), -1 on error. */
int b200_engine_work(b200_engine* e, int timeout_ms);
/* counters: [0] passes that found work while busy-polling, [1] passes that went to epoll_wait,
 * [2] events synthesized by the scan, [3] eventfd (tag) events */
void b200_engine_stats(b200_engine* e, uint64_t out[4]);
/* batching engines: [0] submit calls, [1] Send ops, [2] Recv ops submitted so far */
void b200_engine_batch_stats(b200_engine* e, uint64_t out[3]);
/* Threading: any thread may call b200_endpoint_read / write / shutdown / destroy while another one is inside
 * b200_engine_work (the engine lock is only held while the engine's own state is touched -- never across
 * epoll_wait, a submit, or a user callback).  Callbacks run on the thread that is in b200_engine_work or in
 * the call that completed inline. */

/* grpc_rdma_bp_create, rdma_bp_posix.cc:706-796: takes a pair from the pool, Init, exchanges the
 * 48-byte address over the connected socket `fd` (blocking, exchange_data :640), Connect,
 * registers with the engine and (enable_poller) the background poller.  NULL on failure.
 * The endpoint owns the fd (closed on destroy, like grpc_fd_orphan). */
b200_endpoint* b200_endpoint_create(b200_engine* e, int fd, const char* peer_string, int enable_poller);

/* rdma_read: at most one outstanding.  cb runs from b200_engine_work (or inline when data is
 * already known to be there).  On success the delivered slices are available through
 * b200_endpoint_incoming until the next b200_endpoint_read. */
void b200_endpoint_read(b200_endpoint* ep, b200_closure_fn cb, void* arg, int urgent);
size_t b200_endpoint_incoming(b200_endpoint* ep, const b200_slice** slices);

/* rdma_write: at most one outstanding; the slice array must stay valid until cb runs.
 * A zero-length write completes inline (:566-574). */
void b200_endpoint_write(b200_endpoint* ep, const b200_slice* slices, size_t n, b200_closure_fn cb, void* arg);

/* rdma_shutdown: pending and future callbacks fail with `why`. */
void b200_endpoint_shutdown(b200_endpoint* ep, const char* why);
/* rdma_destroy -> rdma_free :112-132: remove from poller/engine, Disconnect, Putback. */
void b200_endpoint_destroy(b200_endpoint* ep);
const char* b200_endpoint_peer(b200_endpoint* ep);
int b200_endpoint_fd(b200_endpoint* ep);
void* b200_endpoint_pair(b200_endpoint* ep);

/* exchange_data, rdma_bp_posix.cc:640-692: full-duplex exchange of `sz` bytes over a connected
 * socket with poll(2).  0 on success. */
int b200_exchange_data(int fd, const char* buf_in, char* buf_out, size_t sz);

#ifdef __cplusplus
}
#endif
#endif
