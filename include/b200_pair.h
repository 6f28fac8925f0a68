/*
 * b200_pair.h -- C ABI of the B200-native RDMA_BPEV endpoint hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * It replaces, for the reference (pwrliang/grpc-rdma, paths relative to its
 * root), exactly the surface that the endpoint (src/core/lib/iomgr/
 * rdma_bp_posix.cc), the BPEV event engine (src/core/lib/iomgr/
 * ev_epollex_rdma_bpev_linux.cc) and the background Poller (src/core/lib/
 * ibverbs/poller.cc) call on a connection:
 *
 *   grpc_core::ibverbs::PairPollable   src/core/lib/ibverbs/pair.h:82-271
 *   grpc_core::ibverbs::PairPool       src/core/lib/ibverbs/pair.h:273-333
 *   grpc_core::ibverbs::Poller         src/core/lib/ibverbs/poller.h:16-68
 *   grpc_core::ibverbs::Config         src/core/lib/ibverbs/config.h:15-55
 *
 * Implementation: libb200rdma.so (grpc-rdma_b200/csrc).  Ring buffers, credit
 * words and cursors live in HBM; gather/encode (Send), deframe/scatter/clear
 * (Recv), the credit write-back and the readiness scan are sm_100a kernels.
 * There is NO CPU fallback: every data-path entry point fails (returns 0 and
 * sets b200_last_error) if no CUDA device is usable.
 *
 * Conventions kept from the reference: byte counts are returned, never
 * negative; 0 means "nothing moved", the caller then looks at
 * b200_pair_status(); at most one send and one recv may be in flight per pair
 * (ContentAssertion, pair.h:64-81); the has_ and status queries are wait-free
 * and may be called from any thread.
 *
 * Memory rule (the RDMA "registered memory" rule, buffer.cc:9): the batch
 * entry points require slices/destinations that the GPU can address -- device
 * memory, or host memory from b200_mem_alloc_host / b200_mem_register_host.
 * The single-pair entry points accept ANY host pointer; unregistered memory is
 * bounced through a pinned staging buffer (the analogue of send_buffers_
 * [kDataBuffer], pair.cc:104,690).
 */
#ifndef B200_PAIR_H
#define B200_PAIR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_pair b200_pair; /* opaque, pool-owned (PairPollable) */

/* Flattened grpc_slice: GRPC_SLICE_START_PTR / GRPC_SLICE_LENGTH
 * (include/grpc/impl/codegen/slice.h:96-101). */
typedef struct b200_slice {
  const void* ptr;
  uint64_t len;
} b200_slice;

/* PairStatus, pair.h:44-51 (same order, same values). */
enum b200_status {
  B200_UNINITIALIZED = 0,
  B200_INITIALIZED = 1,
  B200_CONNECTED = 2,
  B200_HALF_CLOSED = 3,
  B200_DISCONNECTED = 4,
  B200_ERROR = 5
};

/* Size of the bootstrap blob exchanged over the TCP fd
 * (Address::bytes(), address.h:24-31 / address.cc:19-23; exchange_data,
 * rdma_bp_posix.cc:640-692). */
#define B200_ADDRESS_BYTES 48
/* IBVERBS_PAIR_TAG_POLLABLE, pair.h:26 */
#define B200_PAIR_TAG_POLLABLE 0xa0u
/* GRPC_IBVERBS_POLLER_CAPACITY, poller.h:12 */
#define B200_POLLER_CAPACITY 4096

/* ------------------------------------------------------------------ runtime */

/* Bind the runtime to CUDA device `device` (-1: current device / env
 * B200_DEVICE).  Idempotent.  Returns 0 on success, -1 on failure (no CUDA
 * device, wrong architecture ...); there is no CPU fallback. */
int b200_init(int device);
void b200_shutdown(void);
int b200_device(void);
/* Thread-local description of the last failure ("" if none). */
const char* b200_last_error(void);

/* Config (config.cc:45-115): same keys as the reference's environment
 * variables -- GRPC_RDMA_RING_BUFFER_SIZE_KB (4096), GRPC_RDMA_POLLER_THREAD_NUM
 * (1), GRPC_RDMA_BUSY_POLLING_TIMEOUT_US (500), GRPC_RDMA_POLLER_SLEEP_TIMEOUT_MS
 * (1000), GRPC_RDMA_MAX_SGE (30: what ibv_query_device reported on the authors'
 * HCA, pair.cc:33-35) plus B200_RING_BUFFER_SIZE_BYTES for sub-KB test rings.
 * The environment is read at b200_init; b200_config_set overrides afterwards
 * (affects pairs initialised later).  Returns 0 / -1 (unknown key, bad value). */
int b200_config_set(const char* key, const char* value);
int64_t b200_config_get(const char* key);

/* --------------------------------------------------------------- memory */
void* b200_mem_alloc_device(size_t bytes);
void b200_mem_free_device(void* p);
void* b200_mem_alloc_host(size_t bytes); /* pinned + GPU-addressable (UVA) */
void b200_mem_free_host(void* p);
int b200_mem_register_host(void* p, size_t bytes); /* ibv_reg_mr analogue */
int b200_mem_unregister_host(void* p);
/* Stream-ordered copies (dir: 0 = host->device, 1 = device->host, 2 = d->d). */
int b200_memcpy(void* dst, const void* src, size_t bytes, int dir, void* stream);
int b200_stream_sync(void* stream); /* NULL = the runtime's own stream */

/* ------------------------------------------------------------ pool / pair */

/* PairPool::Take / Putback, pair.h:288-310 */
b200_pair* b200_pool_take(const char* id);
void b200_pool_putback(b200_pair* p);
/* PairPool::Get(id), pair.h:312-320 */
b200_pair* b200_pool_get(const char* id);

/* PairPollable::Init, pair.cc:85-141: (re)allocate + zero the HBM ring and
 * cursors; status -> INITIALIZED. */
void b200_pair_init(b200_pair* p);
/* get_self_address().bytes(), pair.h:150 + address.cc:19: writes
 * B200_ADDRESS_BYTES, returns the size. */
size_t b200_pair_self_address(b200_pair* p, void* out48);
/* PairPollable::Connect, pair.cc:143-168.  1 = connected, 0 = failed (tag or
 * ring size mismatch, peer not reachable by an available wire). */
int b200_pair_connect(b200_pair* p, const void* peer48, size_t n);
/* PairPollable::Disconnect, pair.cc:325-347: tells the peer (peer_exit=1). */
void b200_pair_disconnect(b200_pair* p);

/* PairPollable::Send(grpc_slice*, count, byte_idx), pair.cc:645-734: one frame
 * per slice, <= max_sge frames, a slice is cut only when staging or remote
 * credit runs out.  Returns payload bytes accepted. */
uint64_t b200_pair_send(b200_pair* p, const b200_slice* slices, size_t n, size_t byte_idx);
/* PairPollable::Recv, pair.cc:264-286: at most one frame (or the rest of a
 * partially consumed one) into dst; returns bytes delivered. */
uint64_t b200_pair_recv(b200_pair* p, void* dst, uint64_t cap);

/* pair.cc:288-303 -- wait-free reads of the host-visible mirror that the
 * kernels keep current. */
int b200_pair_has_message(const b200_pair* p);
int b200_pair_has_pending_writes(const b200_pair* p);
uint64_t b200_pair_readable(const b200_pair* p);
uint64_t b200_pair_writable(const b200_pair* p);
/* get_status, pair.cc:349-375; get_error, pair.cc:643 */
enum b200_status b200_pair_status(b200_pair* p);
const char* b200_pair_error(const b200_pair* p);
/* get_wakeup_fd()->read_fd, pair.cc:377: an eventfd the engine registers in
 * epoll with tag ptr|2 (ev_epollex_rdma_bpev_linux.cc:725-741). */
int b200_pair_wakeup_read_fd(b200_pair* p);
/* grpc_wakeup_fd_consume_wakeup on that fd (engine :1018-1021). */
void b200_pair_consume_wakeup(b200_pair* p);

/* Debug / parity inspection (what tests compare with the oracle). */
typedef struct b200_pair_state {
  uint64_t head, moving_head, remain;        /* ring_buffer.h:203-208        */
  uint64_t remote_tail, internal_read_size;  /* pair.h:170-171               */
  uint64_t credit_remote_head;               /* status_report.remote_head    */
  uint32_t partial_write, peer_exit;
  uint64_t ring_capacity;
} b200_pair_state;
int b200_pair_get_state(b200_pair* p, b200_pair_state* out);
/* Copy the pair's HBM ring image to host memory (cap >= ring capacity). */
int b200_pair_copy_ring(b200_pair* p, void* host_dst, uint64_t cap);

/* ------------------------------------------------------------------ poller */

/* Poller::AddPollable / RemovePollable / Shutdown, poller.cc:12-49, poller.h:37.
 * Background thread(s) launch the readiness-scan kernel over all registered
 * pairs and kick a pair's eventfd when it is readable, has a pending partial
 * write, or its peer went away (poller.cc:75-101). */
void b200_poller_add(b200_pair* p);
void b200_poller_remove(b200_pair* p);
void b200_poller_shutdown(void);

/* One synchronous readiness scan over `n` pairs (the body of the engine's
 * busy-poll window, ev_epollex_rdma_bpev_linux.cc:1104-1145): events[i] gets
 * B200_EV_* bits.  Returns the number of pairs with a non-zero event. */
#define B200_EV_READABLE 0x1u /* EPOLLIN: HasMessage, or HalfClosed / Error */
#define B200_EV_WRITABLE 0x4u /* EPOLLOUT: HasPendingWrites */
int b200_poller_scan(b200_pair* const* pairs, size_t n, uint32_t* events);

/* ----------------------------------------------------------------- service */
/*
 * The resident kernels of the unary path (the "persistent warp-per-connection kernel" and the busy-poll half of the
 * BPEV completion loop).  b200_service_start launches three of them and keeps them resident:
 *   owners  one WARP per host command queue.  b200_pair_send / recv post a 128-byte command into the queue of the
 *           pair's connection (both ends of a loopback connection share a queue) and spin on a 16-byte answer.  A
 *           small call (a unary message) is planned, moved, retired and published by that warp alone -- no launch,
 *           no stream synchronisation, no CTA barrier, no lock; a frame that lands at the head of the peer's ring
 *           is pushed to the peer's host slot at once, so the peer's Recv does not need a trip to the GPU.
 *   pool    `workers` CTAs with the k_send / k_recv machinery for everything larger (and for the rdma_flush /
 *           rdma_do_read loops of b200_pairs_submit), fed by the owners through mailboxes in device memory.
 *   poller  scans the connection table continuously, keeps the host-visible mirror of pairs on the nvlink wire
 *           current and appends readiness CHANGES to a ready ring in mapped host memory (one atomic per warp); the
 *           background Poller threads turn those into eventfd kicks instead of launching scans.
 * Returns 0 / -1.  While the service runs: no device-wide synchronisation (cudaDeviceSynchronize, cudaFree; the
 * library defers its own frees until b200_service_stop) and no FIRST launch of a kernel in the process (lazy module
 * loading waits for an idle device; the library loads all of its own kernels before it starts the service).
 */
/* `workers` = pool CTAs (B200_SERVICE_WORKERS, default 16); B200_SERVICE_OWNERS = owner warps = host command queues
 * (default 32).  Fails (-1) when the resident grids would not fit on the device together. */
int b200_service_start(int workers);
/* Call with no b200_pair_send / recv in flight (they would wait for a worker that has left). */
void b200_service_stop(void);
int b200_service_running(void); /* number of pool CTAs, 0 = not running */
/* out[0] commands executed, [1] ready-ring entries consumed, [2] ready-ring overruns,
 * [3] device poller scans (updated every 1024 scans) */
void b200_service_stats(uint64_t out[4]);
/* Recv calls answered from a pair's eagerly pushed host slot (no trip to the GPU and back). */
uint64_t b200_service_eager_hits(void);

/* ------------------------------------------------------------------- batch */
/*
 * B200-native widening of Send/Recv: one kernel launch serves many pairs
 * (what the engine's event loop would otherwise do pair by pair).  Semantics
 * per op are those of the endpoint loops around the single calls:
 *   B200_BATCH_ONE_CALL      exactly one Send / one Recv per op
 *   B200_BATCH_UNTIL_BLOCKED rdma_flush re-entered while Send accepts bytes
 *                            (rdma_bp_posix.cc:470-557) / rdma_do_read's loop
 *                            until dst is full or no complete frame is left
 *                            (rdma_bp_posix.cc:180-286)
 * All pointers must be GPU-addressable (see the memory rule above).
 */
#define B200_BATCH_ONE_CALL 0x0
#define B200_BATCH_UNTIL_BLOCKED 0x1
#define B200_BATCH_ASYNC 0x2 /* do not synchronise; results valid after stream sync */
#define B200_BATCH_ZEROCOPY 0x4 /* pinned HOST buffers are dereferenced by the kernels over PCIe */
/* Sends and Recvs of the SAME connection may run at the same time (batches launched on separate
 * streams without an ordering between them): the kernels then update the host-visible mirrors under a
 * per-pair device lock so that an older readiness / credit view can never overwrite a newer one.
 * The service kernel always works this way; the library's own lanes order the two ends by events and
 * do not need it. */
#define B200_BATCH_CONCURRENT 0x8
/*
 * Where the bytes live decides the path of a batch:
 *   device memory        kernels work in place (one launch per batch);
 *   pinned host memory   HOST-STAGED path (default): the batch owns a device staging arena and
 *                        runs as 8 independent lanes (internal streams), each H2D -> Send kernel
 *                        or Recv kernel -> D2H, so the copy engines and the SMs overlap; a
 *                        connection always maps to the same lane, which keeps its Send and Recv
 *                        ordered.  With a NULL stream the lanes run free (b200_lanes_join or
 *                        b200_batch_results to wait); with a stream the batch forks from / joins
 *                        back into it.  B200_BATCH_ZEROCOPY instead lets the kernels read and
 *                        write the pinned buffers directly (GPUDirect-style, no staging).
 */

/* Threading of this section: a prepared batch belongs to one thread at a time (launch / results / destroy are not
 * locked against each other); different batches may be driven from different threads, the host-staged lanes and the
 * runtime's default stream are shared, so concurrent launches interleave there in submission order.
 * b200_pairs_submit and the post / poll calls are thread-safe (per-queue posting sections, per-thread staging). */
typedef struct b200_send_op {
  b200_pair* pair;
  const b200_slice* slices; /* host array of n entries (copied at submit) */
  size_t nslices;
  size_t byte_idx;
} b200_send_op;

typedef struct b200_recv_op {
  b200_pair* pair;
  void* dst;
  uint64_t cap;
} b200_recv_op;

/* Returns 0 on success.  accepted/delivered: nops entries (may be NULL); with
 * B200_BATCH_ASYNC they must be pinned host memory (b200_mem_alloc_host). */
int b200_pairs_send(const b200_send_op* ops, size_t nops, int flags, uint64_t* accepted, void* stream);
int b200_pairs_recv(const b200_recv_op* ops, size_t nops, int flags, uint64_t* delivered, void* stream);

/* One event-loop pass: all ready Sends and Recvs posted together, waited for together (what
 * pollable_process_events does closure by closure, ev_epollex_rdma_bpev_linux.cc:977-1066).  With the service
 * running nothing is launched: every op is a command of its pair's owner queue, slices (any host memory;
 * unregistered slices are staged) and destinations (GPU-addressable) are used in place.  0 on success. */
int b200_pairs_submit(const b200_send_op* sops, size_t ns, uint64_t* accepted, const b200_recv_op* rops, size_t nr,
                      uint64_t* delivered, int flags);

/* Completion-queue form of the same (service running only): post now, poll later -- the event loop never waits
 * for the GPU, every connection advances at its own pace.  A posted op is an rdma_flush loop / rdma_do_read loop
 * (B200_BATCH_UNTIL_BLOCKED) or a single call.  Recv into pinned HOST memory is delivered into device staging and
 * taken down by the copy engine (one contiguous copy); the op completes when the bytes are in `dst`.
 * post: NULL + *again = 1 when the pair's command queue has no free entry right now (poll something, post later);
 *       NULL + *again = 0 on error.  poll: 1 = finished (bytes valid, handle released), 0 = still running, -1 = error. */
typedef struct b200_async b200_async;
b200_async* b200_pair_post_send(b200_pair* p, const b200_slice* slices, size_t n, size_t byte_idx, int flags, int* again);
b200_async* b200_pair_post_recv(b200_pair* p, void* dst, uint64_t cap, int flags, int* again);
int b200_async_poll(b200_async* op, uint64_t* bytes);

/* Prepared batches: descriptors uploaded to HBM once, launched many times
 * (streaming workloads that reuse their buffers; CUDA-graph friendly). */
typedef struct b200_batch b200_batch;
b200_batch* b200_batch_prepare_send(const b200_send_op* ops, size_t nops, int flags);
b200_batch* b200_batch_prepare_recv(const b200_recv_op* ops, size_t nops, int flags);
int b200_batch_launch(b200_batch* b, void* stream);       /* asynchronous */
/* Per-op byte counts of the most recent launch (synchronises the stream). */
int b200_batch_results(b200_batch* b, uint64_t* out, void* stream);
/* Lanes of the host-staged path: make them wait for `stream` / make `stream` (or, with NULL,
 * the calling thread) wait for them. */
int b200_lanes_fork(void* stream);
int b200_lanes_join(void* stream);
/* Per-op number of Send / Recv calls that moved bytes, as fetched by the last
 * b200_batch_results (parity with the endpoint loops' iteration counts). */
int b200_batch_calls(b200_batch* b, uint64_t* out);
void b200_batch_destroy(b200_batch* b);

/* Calibration: copy `bytes_per_cta` bytes per CTA (CTA i works at offset i*stride of src+mis
 * and dst) with the same one-CTA-per-connection decomposition and copy primitives as the Send
 * kernel but no framing.  Tells what that grid shape can reach on this GPU. */
int b200_probe_copy(void* dst, const void* src, uint64_t bytes_per_cta, uint64_t stride, int nctas, int threads,
                    uint32_t mis, uint32_t item_bytes, uint32_t dynamic, void* stream);

/* Number of kernels this library has launched so far (bench: gpu_launches). */
uint64_t b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_PAIR_H */
