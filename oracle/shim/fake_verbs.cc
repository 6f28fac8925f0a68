// Loopback implementation of the verbs subset declared in
// oracle/shim/infiniband/verbs.h.  TEST INFRASTRUCTURE ONLY (see that header).
//
// One fake device "loop0".  QPs live in a process-wide table keyed by qp_num;
// ibv_modify_qp(RTR) records dest_qp_num.  RDMA_WRITE copies the gather list to
// wr.rdma.remote_addr immediately (same address space) and queues a
// completion; SEND_WITH_IMM is matched in order against the destination QP's
// posted receives (held until one is posted).
#include <infiniband/verbs.h>

#include <deque>
#include <map>
#include <mutex>
#include <vector>

#include <stdlib.h>

namespace {

struct CqImpl {
  std::mutex mu;
  std::deque<ibv_wc> q;
};

struct PendingSend {
  std::vector<uint8_t> bytes;
  uint32_t imm;
};

struct QpImpl {
  uint32_t dest_qpn = 0;
  std::deque<ibv_recv_wr> recvs;     // posted receives (sg_list[0] copied below)
  std::deque<ibv_sge> recv_sges;
  std::deque<PendingSend> inbound;   // sends that arrived before a recv was posted
};

std::mutex g_mu;
std::map<uint32_t, ibv_qp*> g_qps;
uint32_t g_next_qpn = 100;
uint32_t g_next_key = 1000;
ibv_device g_dev;
ibv_device* g_dev_list[2] = {&g_dev, nullptr};

void push_wc(ibv_cq* cq, const ibv_wc& wc) {
  auto* c = static_cast<CqImpl*>(cq->impl);
  std::lock_guard<std::mutex> lk(c->mu);
  c->q.push_back(wc);
}

// g_mu held
void try_match(ibv_qp* dst) {
  auto* d = static_cast<QpImpl*>(dst->impl);
  while (!d->inbound.empty() && !d->recvs.empty()) {
    PendingSend ps = std::move(d->inbound.front());
    d->inbound.pop_front();
    ibv_recv_wr rw = d->recvs.front();
    d->recvs.pop_front();
    ibv_sge sge = d->recv_sges.front();
    d->recv_sges.pop_front();
    size_t n = ps.bytes.size() < sge.length ? ps.bytes.size() : sge.length;
    memcpy(reinterpret_cast<void*>(sge.addr), ps.bytes.data(), n);
    ibv_wc wc;
    memset(&wc, 0, sizeof(wc));
    wc.wr_id = rw.wr_id;
    wc.status = IBV_WC_SUCCESS;
    wc.opcode = IBV_WC_RECV;
    wc.byte_len = static_cast<uint32_t>(n);
    wc.imm_data = ps.imm;
    wc.qp_num = dst->qp_num;
    push_wc(dst->recv_cq, wc);
  }
}

}  // namespace

extern "C" {

ibv_device** ibv_get_device_list(int* num) {
  strcpy(g_dev.name, "loop0");
  if (num) *num = 1;
  return g_dev_list;
}
void ibv_free_device_list(ibv_device**) {}

ibv_context* ibv_open_device(ibv_device* dev) {
  auto* c = new ibv_context;
  c->device = dev;
  return c;
}
int ibv_close_device(ibv_context* ctx) {
  delete ctx;
  return 0;
}
int ibv_query_device(ibv_context*, ibv_device_attr* attr) {
  memset(attr, 0, sizeof(*attr));
  const char* s = getenv("FAKE_VERBS_MAX_SGE");
  attr->max_sge = s ? atoi(s) : 30;  // mlx5 value the authors ran with (rdma_conn.h:21-22)
  attr->max_qp = 1 << 16;
  attr->max_cqe = 1 << 16;
  return 0;
}
int ibv_query_port(ibv_context*, uint8_t, ibv_port_attr* attr) {
  memset(attr, 0, sizeof(*attr));
  attr->state = 4;
  attr->lid = 1;
  attr->link_layer = IBV_LINK_LAYER_INFINIBAND;
  return 0;
}
int ibv_query_gid(ibv_context*, uint8_t, int, ibv_gid* gid) {
  memset(gid, 0, sizeof(*gid));
  return 0;
}
ibv_pd* ibv_alloc_pd(ibv_context* ctx) {
  auto* pd = new ibv_pd;
  pd->context = ctx;
  pd->handle = 1;
  return pd;
}
int ibv_dealloc_pd(ibv_pd* pd) {
  delete pd;
  return 0;
}
ibv_mr* ibv_reg_mr(ibv_pd* pd, void* addr, size_t length, int) {
  auto* mr = new ibv_mr;
  memset(mr, 0, sizeof(*mr));
  mr->context = pd->context;
  mr->pd = pd;
  mr->addr = addr;
  mr->length = length;
  std::lock_guard<std::mutex> lk(g_mu);
  mr->lkey = mr->rkey = g_next_key++;
  return mr;
}
int ibv_dereg_mr(ibv_mr* mr) {
  delete mr;
  return 0;
}
ibv_cq* ibv_create_cq(ibv_context* ctx, int cqe, void* cq_context, ibv_comp_channel*, int) {
  auto* cq = new ibv_cq;
  cq->context = ctx;
  cq->cq_context = cq_context;
  cq->cqe = cqe;
  cq->impl = new CqImpl;
  return cq;
}
int ibv_destroy_cq(ibv_cq* cq) {
  delete static_cast<CqImpl*>(cq->impl);
  delete cq;
  return 0;
}
int ibv_poll_cq(ibv_cq* cq, int n, ibv_wc* wc) {
  auto* c = static_cast<CqImpl*>(cq->impl);
  std::lock_guard<std::mutex> lk(c->mu);
  int got = 0;
  while (got < n && !c->q.empty()) {
    wc[got++] = c->q.front();
    c->q.pop_front();
  }
  return got;
}
ibv_qp* ibv_create_qp(ibv_pd* pd, ibv_qp_init_attr* attr) {
  auto* qp = new ibv_qp;
  memset(qp, 0, sizeof(*qp));
  qp->context = pd->context;
  qp->pd = pd;
  qp->send_cq = attr->send_cq;
  qp->recv_cq = attr->recv_cq;
  qp->qp_type = attr->qp_type;
  qp->state = IBV_QPS_RESET;
  qp->impl = new QpImpl;
  std::lock_guard<std::mutex> lk(g_mu);
  qp->qp_num = g_next_qpn++;
  g_qps[qp->qp_num] = qp;
  return qp;
}
int ibv_destroy_qp(ibv_qp* qp) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_qps.erase(qp->qp_num);
  }
  delete static_cast<QpImpl*>(qp->impl);
  delete qp;
  return 0;
}
int ibv_modify_qp(ibv_qp* qp, ibv_qp_attr* attr, int mask) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto* q = static_cast<QpImpl*>(qp->impl);
  if (mask & IBV_QP_STATE) {
    qp->state = attr->qp_state;
    if (attr->qp_state == IBV_QPS_RESET) {
      q->recvs.clear();
      q->recv_sges.clear();
      q->inbound.clear();
      q->dest_qpn = 0;
    }
  }
  if (mask & IBV_QP_DEST_QPN) q->dest_qpn = attr->dest_qp_num;
  return 0;
}
int ibv_query_qp(ibv_qp* qp, ibv_qp_attr* attr, int, ibv_qp_init_attr*) {
  attr->qp_state = qp->state;
  return 0;
}
int ibv_post_recv(ibv_qp* qp, ibv_recv_wr* wr, ibv_recv_wr** bad) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto* q = static_cast<QpImpl*>(qp->impl);
  for (; wr; wr = wr->next) {
    q->recvs.push_back(*wr);
    q->recv_sges.push_back(wr->sg_list[0]);
  }
  try_match(qp);
  if (bad) *bad = nullptr;
  return 0;
}
int ibv_post_send(ibv_qp* qp, ibv_send_wr* wr, ibv_send_wr** bad) {
  if (bad) *bad = nullptr;
  for (; wr; wr = wr->next) {
    ibv_wc wc;
    memset(&wc, 0, sizeof(wc));
    wc.wr_id = wr->wr_id;
    wc.status = IBV_WC_SUCCESS;
    wc.qp_num = qp->qp_num;
    if (wr->opcode == IBV_WR_RDMA_WRITE) {
      uint8_t* dst = reinterpret_cast<uint8_t*>(wr->wr.rdma.remote_addr);
      for (int i = 0; i < wr->num_sge; i++) {
        // An HCA places an RDMA write in address order; a reader on another thread relies on the
        // last word (a frame's footer, ring_buffer.cc:75-96) landing after everything before it.
        // memcpy gives no such order, so the last 8 bytes of every SGE are stored separately.
        const uint8_t* src = reinterpret_cast<const uint8_t*>(wr->sg_list[i].addr);
        const uint32_t len = wr->sg_list[i].length;
        if (len >= 8 && ((reinterpret_cast<uintptr_t>(dst) + len) & 7) == 0) {
          memcpy(dst, src, len - 8);
          uint64_t last;
          memcpy(&last, src + len - 8, 8);
          __atomic_store_n(reinterpret_cast<uint64_t*>(dst + len - 8), last, __ATOMIC_RELEASE);
        } else {
          memcpy(dst, src, len);
        }
        dst += len;
      }
      wc.opcode = IBV_WC_RDMA_WRITE;
    } else if (wr->opcode == IBV_WR_SEND_WITH_IMM || wr->opcode == IBV_WR_SEND) {
      std::lock_guard<std::mutex> lk(g_mu);
      auto* q = static_cast<QpImpl*>(qp->impl);
      auto it = g_qps.find(q->dest_qpn);
      if (it == g_qps.end()) {
        if (bad) *bad = wr;
        return EINVAL;
      }
      PendingSend ps;
      for (int i = 0; i < wr->num_sge; i++) {
        auto* p = reinterpret_cast<uint8_t*>(wr->sg_list[i].addr);
        ps.bytes.insert(ps.bytes.end(), p, p + wr->sg_list[i].length);
      }
      ps.imm = wr->imm_data;
      static_cast<QpImpl*>(it->second->impl)->inbound.push_back(std::move(ps));
      try_match(it->second);
      wc.opcode = IBV_WC_SEND;
    } else {
      if (bad) *bad = wr;
      return EINVAL;
    }
    if (wr->send_flags & IBV_SEND_SIGNALED) push_wc(qp->send_cq, wc);
  }
  return 0;
}

}  // extern "C"
