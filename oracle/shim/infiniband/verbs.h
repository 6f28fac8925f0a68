/*
 * Loopback stand-in for <infiniband/verbs.h> (rdma-core is not vendored by the
 * reference and there is no NIC in CI).  TEST INFRASTRUCTURE ONLY: it exists so
 * that the reference's src/core/lib/ibverbs/{ring_buffer,pair,poller,device,
 * buffer,memory_region,address,config}.cc compile UNMODIFIED into
 * oracle/_ref/.  Only the types, fields and calls those files touch are
 * declared.  Semantics (fake_verbs.cc): RC in-order placement; RDMA_WRITE is a
 * memcpy to remote_addr (same process), SEND_WITH_IMM matches a posted recv,
 * completions are queued in order.
 */
#ifndef ORACLE_SHIM_INFINIBAND_VERBS_H
#define ORACLE_SHIM_INFINIBAND_VERBS_H
#include <errno.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

union ibv_gid {
  uint8_t raw[16];
  struct {
    uint64_t subnet_prefix;
    uint64_t interface_id;
  } global;
};

struct ibv_device {
  char name[64];
};
struct ibv_context {
  struct ibv_device* device;
};
struct ibv_pd {
  struct ibv_context* context;
  uint32_t handle;
};
struct ibv_mr {
  struct ibv_context* context;
  struct ibv_pd* pd;
  void* addr;
  size_t length;
  uint32_t handle;
  uint32_t lkey;
  uint32_t rkey;
};
struct ibv_comp_channel {
  int fd;
};
struct ibv_cq {
  struct ibv_context* context;
  void* cq_context;
  int cqe;
  void* impl;
};

enum ibv_qp_state { IBV_QPS_RESET, IBV_QPS_INIT, IBV_QPS_RTR, IBV_QPS_RTS, IBV_QPS_SQD, IBV_QPS_SQE, IBV_QPS_ERR };
enum ibv_qp_type { IBV_QPT_RC = 2 };
enum ibv_mtu { IBV_MTU_256 = 1, IBV_MTU_512, IBV_MTU_1024, IBV_MTU_2048, IBV_MTU_4096 };
enum ibv_access_flags { IBV_ACCESS_LOCAL_WRITE = 1, IBV_ACCESS_REMOTE_WRITE = 2, IBV_ACCESS_REMOTE_READ = 4 };
enum ibv_wr_opcode { IBV_WR_RDMA_WRITE, IBV_WR_RDMA_WRITE_WITH_IMM, IBV_WR_SEND, IBV_WR_SEND_WITH_IMM };
enum ibv_send_flags { IBV_SEND_FENCE = 1, IBV_SEND_SIGNALED = 2 };
enum ibv_wc_status { IBV_WC_SUCCESS = 0, IBV_WC_GENERAL_ERR = 1 };
enum ibv_wc_opcode { IBV_WC_SEND, IBV_WC_RDMA_WRITE, IBV_WC_RDMA_READ, IBV_WC_RECV = 128, IBV_WC_RECV_RDMA_WITH_IMM };
enum ibv_qp_attr_mask {
  IBV_QP_STATE = 1 << 0,
  IBV_QP_CUR_STATE = 1 << 1,
  IBV_QP_ACCESS_FLAGS = 1 << 3,
  IBV_QP_PKEY_INDEX = 1 << 4,
  IBV_QP_PORT = 1 << 5,
  IBV_QP_AV = 1 << 7,
  IBV_QP_PATH_MTU = 1 << 8,
  IBV_QP_TIMEOUT = 1 << 9,
  IBV_QP_RETRY_CNT = 1 << 10,
  IBV_QP_RNR_RETRY = 1 << 11,
  IBV_QP_RQ_PSN = 1 << 12,
  IBV_QP_MAX_QP_RD_ATOMIC = 1 << 13,
  IBV_QP_MIN_RNR_TIMER = 1 << 15,
  IBV_QP_SQ_PSN = 1 << 16,
  IBV_QP_MAX_DEST_RD_ATOMIC = 1 << 17,
  IBV_QP_DEST_QPN = 1 << 20
};
enum { IBV_LINK_LAYER_UNSPECIFIED, IBV_LINK_LAYER_INFINIBAND, IBV_LINK_LAYER_ETHERNET };

struct ibv_qp {
  struct ibv_context* context;
  void* qp_context;
  struct ibv_pd* pd;
  struct ibv_cq* send_cq;
  struct ibv_cq* recv_cq;
  uint32_t handle;
  uint32_t qp_num;
  enum ibv_qp_state state;
  enum ibv_qp_type qp_type;
  void* impl;
};

struct ibv_device_attr {
  int max_sge;
  int max_qp;
  int max_cqe;
};
struct ibv_port_attr {
  int state;
  uint16_t lid;
  uint8_t link_layer;
};

struct ibv_qp_cap {
  uint32_t max_send_wr, max_recv_wr, max_send_sge, max_recv_sge, max_inline_data;
};
struct ibv_qp_init_attr {
  void* qp_context;
  struct ibv_cq* send_cq;
  struct ibv_cq* recv_cq;
  void* srq;
  struct ibv_qp_cap cap;
  enum ibv_qp_type qp_type;
  int sq_sig_all;
};
struct ibv_global_route {
  union ibv_gid dgid;
  uint32_t flow_label;
  uint8_t sgid_index, hop_limit, traffic_class;
};
struct ibv_ah_attr {
  struct ibv_global_route grh;
  uint16_t dlid;
  uint8_t sl, src_path_bits, static_rate, is_global, port_num;
};
struct ibv_qp_attr {
  enum ibv_qp_state qp_state, cur_qp_state;
  enum ibv_mtu path_mtu;
  uint32_t qkey, rq_psn, sq_psn, dest_qp_num;
  unsigned int qp_access_flags;
  struct ibv_qp_cap cap;
  struct ibv_ah_attr ah_attr;
  uint16_t pkey_index;
  uint8_t max_rd_atomic, max_dest_rd_atomic, min_rnr_timer, port_num, timeout, retry_cnt, rnr_retry;
};

struct ibv_sge {
  uint64_t addr;
  uint32_t length;
  uint32_t lkey;
};
struct ibv_send_wr {
  uint64_t wr_id;
  struct ibv_send_wr* next;
  struct ibv_sge* sg_list;
  int num_sge;
  enum ibv_wr_opcode opcode;
  unsigned int send_flags;
  uint32_t imm_data;
  union {
    struct {
      uint64_t remote_addr;
      uint32_t rkey;
    } rdma;
  } wr;
};
struct ibv_recv_wr {
  uint64_t wr_id;
  struct ibv_recv_wr* next;
  struct ibv_sge* sg_list;
  int num_sge;
};
struct ibv_wc {
  uint64_t wr_id;
  enum ibv_wc_status status;
  enum ibv_wc_opcode opcode;
  uint32_t vendor_err;
  uint32_t byte_len;
  uint32_t imm_data;
  uint32_t qp_num;
};

struct ibv_device** ibv_get_device_list(int* num);
void ibv_free_device_list(struct ibv_device** list);
struct ibv_context* ibv_open_device(struct ibv_device* dev);
int ibv_close_device(struct ibv_context* ctx);
int ibv_query_device(struct ibv_context* ctx, struct ibv_device_attr* attr);
int ibv_query_port(struct ibv_context* ctx, uint8_t port, struct ibv_port_attr* attr);
int ibv_query_gid(struct ibv_context* ctx, uint8_t port, int index, union ibv_gid* gid);
struct ibv_pd* ibv_alloc_pd(struct ibv_context* ctx);
int ibv_dealloc_pd(struct ibv_pd* pd);
struct ibv_mr* ibv_reg_mr(struct ibv_pd* pd, void* addr, size_t length, int access);
int ibv_dereg_mr(struct ibv_mr* mr);
struct ibv_cq* ibv_create_cq(struct ibv_context* ctx, int cqe, void* cq_context,
                             struct ibv_comp_channel* ch, int vec);
int ibv_destroy_cq(struct ibv_cq* cq);
int ibv_poll_cq(struct ibv_cq* cq, int n, struct ibv_wc* wc);
struct ibv_qp* ibv_create_qp(struct ibv_pd* pd, struct ibv_qp_init_attr* attr);
int ibv_destroy_qp(struct ibv_qp* qp);
int ibv_modify_qp(struct ibv_qp* qp, struct ibv_qp_attr* attr, int mask);
int ibv_query_qp(struct ibv_qp* qp, struct ibv_qp_attr* attr, int mask, struct ibv_qp_init_attr* init);
int ibv_post_send(struct ibv_qp* qp, struct ibv_send_wr* wr, struct ibv_send_wr** bad);
int ibv_post_recv(struct ibv_qp* qp, struct ibv_recv_wr* wr, struct ibv_recv_wr** bad);

#ifdef __cplusplus
}
#endif
#endif
