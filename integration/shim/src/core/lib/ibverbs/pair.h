// pair.h (INTEGRATION SHIM) -- drop-in for the reference's src/core/lib/ibverbs/pair.h.
//
// Same namespace, class names and members the reference's endpoint (src/core/lib/iomgr/
// rdma_bp_posix.cc) and BPEV engine (ev_epollex_rdma_bpev_linux.cc) use today, forwarding to the C ABI
// of include/b200_pair.h.  Put this directory in front of the reference tree on the include path and
// those two files compile UNCHANGED (integration/Makefile `check` does exactly that with
// -fsyntax-only); link with -lb200rdma instead of libibverbs.  Nothing here is used by the product
// library or its tests: it is the binding a maintainer of the reference would add.
#ifndef GRPC_CORE_LIB_IBVERBS_PAIR_H
#define GRPC_CORE_LIB_IBVERBS_PAIR_H

#include <grpc/slice.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "b200_pair.h"
#include "src/core/lib/iomgr/wakeup_fd_posix.h"

namespace grpc_core {
namespace ibverbs {

enum class PairStatus {  // pair.h:44-51, same order as enum b200_status
  kUninitialized,
  kInitialized,
  kConnected,
  kHalfClosed,
  kDisconnected,
  kError
};

class PairPollable {  // pair.h:82-271
 public:
  struct AddressBytes {  // get_self_address().bytes(), pair.h:150 + address.cc:19
    b200_pair* p;
    std::vector<char> bytes() const {
      std::vector<char> v(B200_ADDRESS_BYTES);
      v.resize(b200_pair_self_address(p, v.data()));
      return v;
    }
  };

  explicit PairPollable(b200_pair* p) : p_(p) {
    wakeup_fd_.read_fd = b200_pair_wakeup_read_fd(p);
    wakeup_fd_.write_fd = -1;
  }
  void Init() { b200_pair_init(p_); }                                        // pair.cc:85
  AddressBytes get_self_address() { return AddressBytes{p_}; }
  bool Connect(const std::vector<char>& peer) {                               // pair.cc:143
    return b200_pair_connect(p_, peer.data(), peer.size()) == 1;
  }
  uint64_t Send(grpc_slice* slices, size_t n, size_t byte_idx) {             // pair.cc:645
    // every slice is passed on: a call looks at <= max_sge of them, but the rest counts towards
    // total_slice_size and therefore towards partial_write_ (pair.cc:661-664,712)
    b200_slice stack[64];
    std::vector<b200_slice> heap;
    b200_slice* flat = stack;
    if (n > 64) {
      heap.resize(n);
      flat = heap.data();
    }
    for (size_t i = 0; i < n; i++) flat[i] = {GRPC_SLICE_START_PTR(slices[i]), GRPC_SLICE_LENGTH(slices[i])};
    return b200_pair_send(p_, flat, n, byte_idx);
  }
  uint64_t Recv(void* buf, uint64_t cap) { return b200_pair_recv(p_, buf, cap); }   // pair.cc:264
  bool HasMessage() const { return b200_pair_has_message(p_) != 0; }                // pair.cc:288
  bool HasPendingWrites() const { return b200_pair_has_pending_writes(p_) != 0; }   // pair.cc:303
  uint64_t GetReadableSize() const { return b200_pair_readable(p_); }               // pair.cc:290
  uint64_t GetWritableSize() const { return b200_pair_writable(p_); }               // pair.cc:294
  PairStatus get_status() { return static_cast<PairStatus>(b200_pair_status(p_)); } // pair.cc:349
  const std::string& get_error() {                                                  // pair.cc:643
    err_ = b200_pair_error(p_);
    return err_;
  }
  grpc_wakeup_fd* get_wakeup_fd() { return &wakeup_fd_; }                           // pair.cc:377
  void Disconnect() { b200_pair_disconnect(p_); }                                   // pair.cc:325
  b200_pair* raw() { return p_; }

 private:
  b200_pair* p_;
  grpc_wakeup_fd wakeup_fd_;
  std::string err_;
};

class PairPool {  // pair.h:273-333
 public:
  static PairPool& Get() {
    static PairPool pool;
    return pool;
  }
  PairPollable* Take(const std::string& id) {
    b200_pair* p = b200_pool_take(id.c_str());
    return p ? new PairPollable(p) : nullptr;
  }
  void Putback(PairPollable* p) {
    b200_pool_putback(p->raw());
    delete p;
  }
};

}  // namespace ibverbs
}  // namespace grpc_core
#endif
