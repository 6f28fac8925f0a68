// poller.h (INTEGRATION SHIM) -- drop-in for the reference's src/core/lib/ibverbs/poller.h (:16-68).
#ifndef GRPC_CORE_LIB_IBVERBS_POLLER_H
#define GRPC_CORE_LIB_IBVERBS_POLLER_H
#include "src/core/lib/ibverbs/pair.h"

namespace grpc_core {
namespace ibverbs {
class Poller {
 public:
  static Poller& Get() {
    static Poller p;
    return p;
  }
  void AddPollable(PairPollable* p) { b200_poller_add(p->raw()); }        // poller.cc:12
  void RemovePollable(PairPollable* p) { b200_poller_remove(p->raw()); }  // poller.cc:41
  void Shutdown() { b200_poller_shutdown(); }                             // poller.h:37
};
}  // namespace ibverbs
}  // namespace grpc_core
#endif
