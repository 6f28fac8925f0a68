/* infiniband/verbs.h for the offline build of the reference's full stack (integration/stack).
 * = the loopback stand-in of oracle/shim (installed next to this file as verbs_base.h by
 * build_stack.sh) + the declarations that only the reference's RDMA_EVENT mode touches
 * (src/core/lib/rdma/ *: completion channels, async events, device names).  That mode is out of
 * scope; its entry points are link-only stubs in verbs_event_stubs.cc. */
#ifndef STACK_INFINIBAND_VERBS_H
#define STACK_INFINIBAND_VERBS_H
#include "verbs_base.h"
#ifdef __cplusplus
extern "C" {
#endif
struct ibv_srq;
enum ibv_event_type {
  IBV_EVENT_CQ_ERR, IBV_EVENT_QP_FATAL, IBV_EVENT_QP_REQ_ERR, IBV_EVENT_QP_ACCESS_ERR, IBV_EVENT_COMM_EST,
  IBV_EVENT_SQ_DRAINED, IBV_EVENT_PATH_MIG, IBV_EVENT_PATH_MIG_ERR, IBV_EVENT_DEVICE_FATAL, IBV_EVENT_PORT_ACTIVE,
  IBV_EVENT_PORT_ERR, IBV_EVENT_LID_CHANGE, IBV_EVENT_PKEY_CHANGE, IBV_EVENT_SM_CHANGE, IBV_EVENT_SRQ_ERR,
  IBV_EVENT_SRQ_LIMIT_REACHED, IBV_EVENT_QP_LAST_WQE_REACHED, IBV_EVENT_CLIENT_REREGISTER, IBV_EVENT_GID_CHANGE
};
struct ibv_async_event {
  union {
    struct ibv_cq* cq;
    struct ibv_qp* qp;
    struct ibv_srq* srq;
    int port_num;
  } element;
  enum ibv_event_type event_type;
};
const char* ibv_get_device_name(struct ibv_device* dev);
struct ibv_comp_channel* ibv_create_comp_channel(struct ibv_context* ctx);
int ibv_destroy_comp_channel(struct ibv_comp_channel* ch);
int ibv_req_notify_cq(struct ibv_cq* cq, int solicited_only);
int ibv_get_cq_event(struct ibv_comp_channel* ch, struct ibv_cq** cq, void** cq_context);
void ibv_ack_cq_events(struct ibv_cq* cq, unsigned int nevents);
#ifdef __cplusplus
}
#endif
#endif
