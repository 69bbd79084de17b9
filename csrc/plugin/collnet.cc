// Collective offload for NCCL: ncclCollNetPlugin_v4 / _v6 / _v7 / _v8 (+ _v9 / _v10 in the -bnetx build) — iallreduce runs the
// two-shot all-reduce of csrc/coll/transport_mesh.cc over a full mesh of this plugin's own connections: the reduction is
// done by the SENDING kernels while they move the data over NVLink (fused isend, K4 of SURVEY.md section 2.6), no kernel
// of NCCL touches the payload, the result lands in NCCL's receive buffer.
//
// The reference carries the CollNet declaration only (reference cc/v4/nccl_net_v4.h:64-101) and exports no
// ncclCollNetPlugin symbol (SURVEY.md section 2.5 "collective algorithms: absent"; section 7.2 step 6 "CollNet export
// experiment").  What NCCL does with the table:
//     init -> devices -> getProperties -> listen (every rank) -> [NCCL all-gathers the handles] -> connect(handles, nranks, rank)
//     -> reduceSupport -> regMr(send buffer), regMr(receive buffer) -> iallreduce / test ... -> closeColl, closeListen
// and only when the user asks for it (NCCL_COLLNET_ENABLE=1; one rank per "node", e.g. NCCL_HOSTID per rank, or
// NCCL_COLLNET_NODE_THRESHOLD).  BNET_COLLNET=1 makes the table report its devices; without it `devices` answers 0 and NCCL
// drops the table at init — the net path of the same library is untouched either way.
//
// Semantics:
//   * sum of fp32 or bf16 (reduceSupport says so); everything else is left to NCCL's own algorithms
//   * device memory (NVLink transport: the reduction is fused into the isends) or host memory (any transport — multi-stream
//     TCP between hosts, what NCCL hands a plugin without GPUDirect: the reduce-scatter pieces are added on the host);
//     the send and the receive buffer of one call must be of the same kind
//   * one all-reduce at a time per collComm, in call order (the mesh matches messages strictly FIFO per connection);
//     up to 8 further calls queue behind it, beyond that iallreduce returns request = NULL ("try again", like isend)
//   * iflush: the sender's kernel fences at system scope before the completion word is visible, nothing is left to flush
//   * iallgather / ireducescatter (v8+): not offered (ncclInvalidUsage); NCCL only calls them for its NVLS + CollNet mode
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <deque>
#include <vector>

#include "bnet/nccl_net_abi.h"
#include "coll/transport_mesh.h"
#include "core/engine.h"
#include "plugin/plugin_shims.h"

using namespace bnet;

#define BNET_EXPORT __attribute__((visibility("default")))
#pragma GCC diagnostic ignored "-Wunused-function"   // (the v9 / v10 entry points are only referenced in the -bnetx build)

namespace {

const char kCollName[] = "BNet";
constexpr size_t kMaxQueued = 8;

struct CollListen {
  ListenComm* listen = nullptr;
  int dev = 0;
};

struct CollComm;
struct CollReq {
  CollComm* comm = nullptr;
  bool flush = false;
  const void* send = nullptr;
  void* recv = nullptr;
  size_t count = 0;
  int dtype = 0;                 // 0 = fp32, 1 = bf16
  MeshMr* smr = nullptr;
  MeshMr* rmr = nullptr;
  MeshOp* op = nullptr;          // created when the request reaches the head of the queue
  int state = 0;                 // 0 = queued, 1 = running, 2 = finished, -1 = failed
};

struct CollComm {
  BnetTMesh* mesh = nullptr;
  int rank = 0, nranks = 0;
  std::deque<CollReq*> queue;    // head = the all-reduce that is running
  size_t piece_bytes = 1 << 20;
  int inflight = 8;
  int timeout_ms = 0;
};

std::atomic<uint64_t> g_allreduces_done{0};     // all-reduces this process completed through the table (bnet_collnet_allreduces)

bool collnet_enabled() {
  static const bool on = env_int("COLLNET", 0) != 0;
  return on;
}

ncclResult_t coll_init(ncclDebugLogger_t logfn) { return plugin::init(logfn); }

ncclResult_t coll_devices(int* ndev) {
  if (!ndev) return ncclInvalidArgument;
  *ndev = collnet_enabled() ? Engine::get().ndev() : 0;
  return ncclSuccess;
}

template <typename P, ncclResult_t (*F)(int, P*)>
ncclResult_t coll_props(int dev, P* o) {
  ncclResult_t r = F(dev, o);
  // device memory rides the fused isends of the NVLink transport; host memory is reduced over any transport (TCP between hosts)
  if (r == ncclSuccess) o->ptrSupport = NCCL_PTR_HOST | (Engine::get().cuda_ok() ? NCCL_PTR_CUDA : 0);
  return r;
}

ncclResult_t coll_listen(int dev, void* handle, void** lcomm) {
  if (!handle || !lcomm) return ncclInvalidArgument;
  ListenComm* l = nullptr;
  int st = Engine::get().listen(dev, handle, NCCL_NET_HANDLE_MAXSIZE, &l);
  if (st) {
    BNET_WARN("collnet listen(dev=%d) failed: %s", dev, status_str(st));
    return to_nccl(st);
  }
  CollListen* cl = new CollListen();
  cl->listen = l;
  cl->dev = dev;
  *lcomm = cl;
  return ncclSuccess;
}

ncclResult_t coll_listen_v4(int dev, void* handle, void** lcomm) {
  // (the v4 handle is 64 bytes: the engine's Handle fits, and connect() below reads 64-byte entries for that table)
  if (!handle || !lcomm) return ncclInvalidArgument;
  ListenComm* l = nullptr;
  int st = Engine::get().listen(dev, handle, NCCL_NET_HANDLE_MAXSIZE_V4, &l);
  if (st) return to_nccl(st);
  CollListen* cl = new CollListen();
  cl->listen = l;
  cl->dev = dev;
  *lcomm = cl;
  return ncclSuccess;
}

ncclResult_t connect_sized(void* handles[], int nranks, int rank, void* lcomm, void** ccomm, size_t handle_bytes) {
  if (!handles || !lcomm || !ccomm || nranks < 2 || nranks > 64 || rank < 0 || rank >= nranks) return ncclInvalidArgument;
  CollListen* cl = static_cast<CollListen*>(lcomm);
  if (!cl->listen) return ncclInvalidUsage;      // one collComm per listenComm
  std::vector<char> blob((size_t)nranks * NCCL_NET_HANDLE_MAXSIZE, 0);
  for (int r = 0; r < nranks; r++) {
    if (!handles[r]) return ncclInvalidArgument;
    memcpy(blob.data() + (size_t)r * NCCL_NET_HANDLE_MAXSIZE, handles[r], handle_bytes);
  }
  BnetTMesh* m = tmesh_new(cl->listen, rank, nranks, cl->dev);
  if (!m) return ncclInternalError;
  cl->listen = nullptr;                          // the mesh owns the listening socket from here on
  const int timeout_ms = (int)env_int("COLLNET_CONNECT_TIMEOUT_MS", 60000);
  if (tmesh_connect(m, blob.data(), timeout_ms) != 0) {
    BNET_WARN("collnet connect (rank %d of %d) failed: %s", rank, nranks, tmesh_error(m));
    tmesh_destroy(m);
    return ncclSystemError;
  }
  CollComm* c = new CollComm();
  c->mesh = m;
  c->rank = rank;
  c->nranks = nranks;
  c->piece_bytes = (size_t)env_int("COLLNET_PIECE_BYTES", 1 << 20);
  c->inflight = (int)env_int("COLLNET_INFLIGHT", 8);
  c->timeout_ms = (int)env_int("TIMEOUT_MS", 0);
  BNET_INFO("collnet: rank %d of %d connected (two-shot all-reduce over %d + %d connections)", rank, nranks, nranks - 1, nranks - 1);
  *ccomm = c;
  return ncclSuccess;
}
ncclResult_t coll_connect(void* handles[], int nranks, int rank, void* lcomm, void** ccomm) {
  return connect_sized(handles, nranks, rank, lcomm, ccomm, NCCL_NET_HANDLE_MAXSIZE);
}
ncclResult_t coll_connect_v4(void* handles[], int nranks, int rank, void* lcomm, void** ccomm) {
  return connect_sized(handles, nranks, rank, lcomm, ccomm, NCCL_NET_HANDLE_MAXSIZE_V4);
}

ncclResult_t coll_reduce_support(ncclDataType_t dt, ncclRedOp_t op, int* supported) {
  if (!supported) return ncclInvalidArgument;
  *supported = (op == ncclSum && (dt == ncclFloat32 || dt == ncclBfloat16)) ? 1 : 0;
  return ncclSuccess;
}

ncclResult_t coll_regmr(void* ccomm, void* data, size_t size, int type, void** mhandle) {
  if (!ccomm || !mhandle) return ncclInvalidArgument;
  CollComm* c = static_cast<CollComm*>(ccomm);
  if (type != NCCL_PTR_CUDA && type != NCCL_PTR_HOST) {
    BNET_WARN("collnet regMr: unsupported pointer type %d", type);
    return ncclInvalidUsage;
  }
  MeshMr* mr = tmesh_reg(c->mesh, data, size, type);
  if (!mr) {
    BNET_WARN("collnet regMr(%p, %zu) failed: %s", data, size, tmesh_error(c->mesh));
    return ncclSystemError;
  }
  *mhandle = mr;
  return ncclSuccess;
}
ncclResult_t coll_regmr_int(void* ccomm, void* data, int size, int type, void** mhandle) {
  return coll_regmr(ccomm, data, (size_t)size, type, mhandle);
}
ncclResult_t coll_regmr_dmabuf(void* ccomm, void* data, size_t size, int type, uint64_t, int, void** mhandle) {
  return coll_regmr(ccomm, data, size, type, mhandle);   // NCCL_PTR_DMABUF is never advertised
}

ncclResult_t coll_deregmr(void* ccomm, void* mhandle) {
  if (!ccomm) return ncclInvalidArgument;
  if (mhandle) tmesh_dereg(static_cast<CollComm*>(ccomm)->mesh, static_cast<MeshMr*>(mhandle));
  return ncclSuccess;
}

// Start whatever stands at the head of the queue.  A request that cannot start is marked failed and leaves the queue (its
// owner sees the error in its own test()); the next one gets its turn.
void advance(CollComm* c) {
  while (!c->queue.empty() && c->queue.front()->state == 0) {
    CollReq* q = c->queue.front();
    q->op = tmesh_op_start(c->mesh, MESH_TWO_SHOT, q->send, q->smr, q->recv, q->rmr, q->count, q->dtype, q->dtype, c->piece_bytes,
                           c->inflight, c->timeout_ms);
    if (q->op) {
      q->state = 1;
      return;
    }
    BNET_WARN("collnet iallreduce(%zu elements) could not start: %s", q->count, tmesh_error(c->mesh));
    q->state = -1;
    c->queue.pop_front();
  }
}

ncclResult_t coll_iallreduce(void* ccomm, void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, void* smh,
                             void* rmh, void** request) {
  if (!ccomm || !request || !smh || !rmh) return ncclInvalidArgument;
  CollComm* c = static_cast<CollComm*>(ccomm);
  int ok = 0;
  coll_reduce_support(dt, op, &ok);
  if (!ok) return ncclInvalidUsage;
  *request = nullptr;
  if (c->queue.size() >= kMaxQueued + 1) return ncclSuccess;    // request stays NULL: try again later
  CollReq* q = new CollReq();
  q->comm = c;
  q->send = send;
  q->recv = recv;
  q->count = count;
  q->dtype = dt == ncclFloat32 ? 0 : 1;
  q->smr = static_cast<MeshMr*>(smh);
  q->rmr = static_cast<MeshMr*>(rmh);
  c->queue.push_back(q);
  advance(c);
  if (q->state == -1) {            // (advance() has taken it out of the queue)
    delete q;
    return ncclInternalError;
  }
  *request = q;
  return ncclSuccess;
}
ncclResult_t coll_iallreduce_int(void* ccomm, void* send, void* recv, int count, ncclDataType_t dt, ncclRedOp_t op, void* smh,
                                 void* rmh, void** request) {
  if (count < 0) return ncclInvalidArgument;
  return coll_iallreduce(ccomm, send, recv, (size_t)count, dt, op, smh, rmh, request);
}

ncclResult_t coll_iallgather_v8(void*, void*, int, ncclNetSGE_v8_t*, size_t, size_t, size_t, void*, void**) {
  BNET_WARN("collnet: iallgather is not offered");
  return ncclInvalidUsage;
}
ncclResult_t coll_ireducescatter_v8(void*, int, ncclNetSGE_v8_t*, void*, size_t, size_t, size_t, ncclDataType_t, ncclRedOp_t, void*,
                                    void**) {
  BNET_WARN("collnet: ireducescatter is not offered");
  return ncclInvalidUsage;
}
ncclResult_t coll_iallgather_v9(void*, void*, int, ncclNetSGE_v9_t*, size_t, size_t, size_t, void*, void**) {
  BNET_WARN("collnet: iallgather is not offered");
  return ncclInvalidUsage;
}
ncclResult_t coll_ireducescatter_v9(void*, int, ncclNetSGE_v9_t*, void*, size_t, size_t, size_t, ncclDataType_t, ncclRedOp_t, void*,
                                    void**) {
  BNET_WARN("collnet: ireducescatter is not offered");
  return ncclInvalidUsage;
}

ncclResult_t coll_iflush(void* ccomm, void*, int, void*, void** request) {
  if (!ccomm || !request) return ncclInvalidArgument;
  CollReq* q = new CollReq();
  q->comm = static_cast<CollComm*>(ccomm);
  q->flush = true;
  q->state = 2;
  *request = q;
  return ncclSuccess;
}

ncclResult_t coll_test(void* request, int* done, int* size) {
  if (!request || !done) return ncclInvalidArgument;
  CollReq* q = static_cast<CollReq*>(request);
  CollComm* c = q->comm;
  *done = 0;
  if (q->flush) {
    *done = 1;
    if (size) *size = 0;
    delete q;
    return ncclSuccess;
  }
  // drive the all-reduce at the head of the queue (which may or may not be this request)
  if (!c->queue.empty() && c->queue.front()->state == 1) {
    CollReq* h = c->queue.front();
    const int st = tmesh_op_step(h->op);
    if (st != 0) {
      if (st < 0) BNET_WARN("collnet all-reduce failed: %s", tmesh_error(c->mesh));
      tmesh_op_free(h->op);
      h->op = nullptr;
      h->state = st > 0 ? 2 : -1;
      if (st > 0) g_allreduces_done.fetch_add(1, std::memory_order_relaxed);
      c->queue.pop_front();        // (the request object lives until its owner has tested it)
      advance(c);
    }
  }
  if (q->state == -1) {
    delete q;
    return ncclInternalError;
  }
  if (q->state == 2) {
    *done = 1;
    if (size) *size = (int)(q->count * (q->dtype == 0 ? 4 : 2));
    delete q;
  }
  return ncclSuccess;
}

ncclResult_t coll_close(void* ccomm) {
  CollComm* c = static_cast<CollComm*>(ccomm);
  if (!c) return ncclSuccess;
  for (CollReq* q : c->queue) {     // NCCL closes a comm after its requests have completed; whatever is left is dropped
    if (q->op) tmesh_op_free(q->op);
    delete q;
  }
  tmesh_destroy(c->mesh);
  delete c;
  return ncclSuccess;
}

ncclResult_t coll_close_listen(void* lcomm) {
  CollListen* cl = static_cast<CollListen*>(lcomm);
  if (!cl) return ncclSuccess;
  delete cl->listen;     // nullptr once a collComm has adopted it
  delete cl;
  return ncclSuccess;
}

ncclResult_t coll_make_vdevice(int*, ncclNetVDeviceProps_v9_t*) { return ncclInvalidUsage; }   // no NIC fusion

}  // namespace

extern "C" {
// evidence for "did NCCL really run its all-reduces through the table": completed iallreduce calls of this process
BNET_EXPORT unsigned long long bnet_collnet_allreduces(void) { return g_allreduces_done.load(std::memory_order_relaxed); }

BNET_EXPORT ncclCollNet_v4_t ncclCollNetPlugin_v4 = {
    kCollName, coll_init, coll_devices, coll_props<ncclNetProperties_v4_t, plugin::v4_props>, coll_listen_v4, coll_connect_v4,
    coll_reduce_support, coll_regmr_int, coll_deregmr, coll_iallreduce_int, coll_iflush, coll_test, coll_close, coll_close_listen};

BNET_EXPORT ncclCollNet_v6_t ncclCollNetPlugin_v6 = {
    kCollName, coll_init, coll_devices, coll_props<ncclNetProperties_v6_t, plugin::v6_props>, coll_listen, coll_connect,
    coll_reduce_support, coll_regmr_int, coll_regmr_dmabuf, coll_deregmr, coll_iallreduce_int, coll_iflush, coll_test, coll_close,
    coll_close_listen};

BNET_EXPORT ncclCollNet_v7_t ncclCollNetPlugin_v7 = {
    kCollName, coll_init, coll_devices, coll_props<ncclNetProperties_v7_t, plugin::v7_props>, coll_listen, coll_connect,
    coll_reduce_support, coll_regmr_int, coll_regmr_dmabuf, coll_deregmr, coll_iallreduce_int, coll_iflush, coll_test, coll_close,
    coll_close_listen};

BNET_EXPORT ncclCollNet_v8_t ncclCollNetPlugin_v8 = {
    kCollName, coll_init, coll_devices, coll_props<ncclNetProperties_v8_t, plugin::v8_props>, coll_listen, coll_connect,
    coll_reduce_support, coll_regmr, coll_regmr_dmabuf, coll_deregmr, coll_iallreduce_int, coll_iallgather_v8, coll_ireducescatter_v8,
    coll_iflush, coll_test, coll_close, coll_close_listen};

#ifdef BNET_EXPORT_V9_V10
BNET_EXPORT ncclCollNet_v9_t ncclCollNetPlugin_v9 = {
    kCollName, coll_init, coll_devices, coll_props<ncclNetProperties_v9_t, plugin::v9_props>, coll_listen, coll_connect,
    coll_reduce_support, coll_regmr, coll_regmr_dmabuf, coll_deregmr, coll_iallreduce, coll_iallgather_v9, coll_ireducescatter_v9,
    coll_iflush, coll_test, coll_close, coll_close_listen, coll_make_vdevice};

BNET_EXPORT ncclCollNet_v10_t ncclCollNetPlugin_v10 = {
    kCollName, coll_init, coll_devices, coll_props<ncclNetProperties_v9_t, plugin::v9_props>, coll_listen, coll_connect,
    coll_reduce_support, coll_regmr, coll_regmr_dmabuf, coll_deregmr, coll_iallreduce, coll_iallgather_v9, coll_ireducescatter_v9,
    coll_iflush, coll_test, coll_close, coll_close_listen, coll_make_vdevice};
#endif
}  // extern "C"
