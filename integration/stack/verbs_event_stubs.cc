// Link-only stubs for the verbs entry points that only the reference's RDMA_EVENT mode calls
// (src/core/lib/rdma/*; out of scope).  They report "not supported".
#include <infiniband/verbs.h>
#include <errno.h>
extern "C" {
const char* ibv_get_device_name(struct ibv_device* dev) { return dev ? dev->name : "loop0"; }
struct ibv_comp_channel* ibv_create_comp_channel(struct ibv_context*) { errno = ENOSYS; return nullptr; }
int ibv_destroy_comp_channel(struct ibv_comp_channel*) { return 0; }
int ibv_req_notify_cq(struct ibv_cq*, int) { return ENOSYS; }
int ibv_get_cq_event(struct ibv_comp_channel*, struct ibv_cq**, void**) { return -1; }
void ibv_ack_cq_events(struct ibv_cq*, unsigned int) {}
}
