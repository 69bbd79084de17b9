// All-reduce that RIDES THE TRANSPORT over a full mesh of plugin connections: the sending kernel accumulates into the peer's
// buffer while it moves the data (fused isend, K4 / K4+K5 of SURVEY.md section 2.6), so no reduce kernel and no staging
// buffer exist anywhere.  Two shapes, both built from the same resumable operation (MeshOp):
//
//   ONE-SHOT  every rank sends its whole input to every peer once.  One network step, (n-1) x size on the wire per rank:
//             the latency-optimal shape up to a few MiB.
//               out <- in                      local executor pass (copy, or bf16 -> fp32 cast when the output is wider)
//               irecv(out) on every incoming connection            (posted only after the local pass)
//               isend_op(in, add) on every outgoing connection — n-1 kernels accumulate into each output concurrently
//             The n-1 senders race, so ranks agree bit for bit only when the additions are exact.
//
//   TWO-SHOT  the buffer is cut into n slices, rank r owns slice r.
//               A (reduce-scatter): out[slice r] <- in[slice r] locally; every peer q gets isend_op(in[slice q], add) —
//                                   n-1 kernels accumulate into the owner's slice
//               B (all-gather)    : once every contribution has ARRIVED, the owner copies its finished slice to every peer
//             Two network steps whatever n is (the ring of transport_ring.cc needs 2(n-1)), 2(n-1)/n x size on the wire per
//             rank — what a switch wants: every link of the NVSwitch fabric busy in both phases.  Every rank ends with the
//             owner's bits, so ranks agree exactly even for floats (the order INSIDE a slice's sum is still not fixed).
//             In-place (in == out) is allowed.  This is the algorithm behind the CollNet table (csrc/plugin/collnet.cc).
//
//   HOST MODE  buffers registered as host memory (NCCL_PTR_HOST) are reduced with the two-shot schedule over ANY transport —
//             multi-stream TCP between hosts, the same-host shared-memory ring: phase A pieces are plain isends into a
//             per-connection staging area of the owner, which adds them into its slice on the host when they have arrived
//             (no fused isend exists off NVLink); phase B is the same copy.  This is what the CollNet table runs when NCCL
//             hands it host buffers (no GPUDirect), i.e. the all-reduce offload for the reference's own deployment.
//
// Messages are cut into pieces, `inflight` requests per connection and direction; strictly FIFO per connection (the
// transport's matching rule), phase A before phase B.  The incoming connections carry no identity of their own, so the
// first message on every connection is the sender's rank (4 bytes of host memory).
//
// The reference has no collectives (reference README.md:88; its ncclCollNet_v4_t is a declaration
// only, reference cc/v4/nccl_net_v4.h:64-101); this and the ring are what section 7.2 step 4 of the survey asks for.
#include "coll/backoff.h"
#include "coll/transport_mesh.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "core/telemetry.h"
#include "cuda/cuda_iface.h"
#include "cuda/exec_ops.h"
#include "cuda/nvl_exec.h"

using namespace bnet;

#define BNET_API extern "C" __attribute__((visibility("default")))

struct BnetTMesh {
  int rank = 0, world = 1, dev = 0;
  ListenComm* listen = nullptr;
  std::vector<Comm*> send;          // world - 1 outgoing connections; send[c] goes to rank (rank + 1 + c) % world
  std::vector<Comm*> recv;          // world - 1 incoming connections in arrival order; recv[c] comes from recv_peer[c]
  std::vector<int> recv_peer;
  bool identified = false;
  MeshMr* in_mr = nullptr;          // the C API's registered pair (bnet_tmesh_register)
  MeshMr* out_mr = nullptr;
  volatile uint64_t* jflags = nullptr;   // completion words of the local pass
  uint64_t* jflags_dev = nullptr;
  uint64_t jseq = 0;
  uint64_t op_seq = 0;              // all-reduces started on this mesh (span ids)
  MeshOp* active = nullptr;         // one operation at a time (FIFO matching per connection)
  std::vector<std::vector<char>> staging;   // host mode: per incoming connection, where phase-A pieces land (kept between calls:
                                            // a fresh 8 MiB vector per connection and call costs a memset and its page faults)
  char err[256] = {0};
  uint64_t last_msgs = 0, last_bytes_sent = 0;
};

namespace bnet {

struct MeshMr {
  char* base = nullptr;
  size_t bytes = 0;
  int type = NCCL_PTR_CUDA;
  std::vector<MemHandle*> smh, rmh;   // per outgoing / incoming connection
};

namespace {

int fail(BnetTMesh* m, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(BnetTMesh* m, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(m->err, sizeof(m->err), fmt, ap);
  va_end(ap);
  BNET_WARN("transport mesh: %s", m->err);
  return -1;
}

struct MeshMsg {
  char* ptr;
  size_t len;
  uint32_t op;        // sends: ExecOp (0 = plain isend)
  MemHandle* mh;
  uint8_t phase;      // 0 = A, 1 = B (B sends wait for every A receive)
  char* acc = nullptr;   // host mode, phase A receives: the piece landed in a staging area and is added into this place
};

// out[i] += in[i] on the host (the reduce-scatter of the host mode); dtype pairs as everywhere in this file
__attribute__((optimize("O3"))) void host_accumulate(char* out, const char* in, size_t in_bytes, int in_dtype, int out_dtype) {
  if (in_dtype == 0) {
    float* __restrict__ o = reinterpret_cast<float*>(out);            // (the staging area never overlaps the output)
    const float* __restrict__ a = reinterpret_cast<const float*>(in);
    for (size_t i = 0, n = in_bytes / 4; i < n; i++) o[i] += a[i];
    return;
  }
  const uint16_t* a = reinterpret_cast<const uint16_t*>(in);
  auto widen = [](uint16_t v) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); return f; };
  const size_t n = in_bytes / 2;
  if (out_dtype == 0) {
    float* o = reinterpret_cast<float*>(out);
    for (size_t i = 0; i < n; i++) o[i] += widen(a[i]);
    return;
  }
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  for (size_t i = 0; i < n; i++) {                     // bf16 += bf16: add in fp32, round to nearest even (like the device op)
    float f = widen(o[i]) + widen(a[i]);
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { o[i] = (uint16_t)((u >> 16) | 0x40); continue; }
    u += 0x7fffu + ((u >> 16) & 1);
    o[i] = (uint16_t)(u >> 16);
  }
}

struct MeshSide {
  std::vector<MeshMsg> msgs;
  std::vector<Request*> req;
  std::vector<unsigned char> fin;
  size_t posted = 0, done = 0;
};

}  // namespace

struct MeshOp {
  BnetTMesh* m = nullptr;
  std::vector<MeshSide> rs, ss;
  size_t remaining = 0;       // requests not finished yet (posted or not)
  size_t a_recv_left = 0;     // phase A receives not finished yet
  int stage = 0;              // 0 = local pass running, 1 = exchanging, 2 = finished, -1 = failed
  uint64_t init_value = 0;
  int init_chunks = 0;
  int cuda_dev = 0;
  int inflight = 8;
  int timeout_ms = 0;
  uint64_t t0 = 0;
  uint64_t span = 0;          // "coll-<rank>" span around the whole operation (its isend / irecv spans fall inside it)
  uint64_t out_bytes = 0;
  // host mode (buffers in host memory, any transport): no fused isend exists, so phase A pieces land in per-connection
  // staging areas and are added into the own slice here, on the host; phase B is unchanged
  bool host = false;
  int in_dtype = 0, out_dtype = 0;
  unsigned idle = 0;          // consecutive steps in which nothing moved (the blocking wrapper backs off on it)
  // In place, over transports whose sends are served by other threads (TCP): the reduce-scatter sends to connection c's peer
  // that have not completed locally.  That peer's finished slice arrives in the memory those sends read — only after it got
  // them, by causality through the network; the receive is nevertheless posted only when the count is zero (an explicit
  // happens-before).  Empty = not tracked: out of place, or the NVLink path (kernels on both ends), which posts ahead.
  std::vector<uint32_t> a_send_left;
};

BnetTMesh* tmesh_new(ListenComm* listen, int rank, int world, int net_dev) {
  if (!listen || world < 2 || world > 64 || rank < 0 || rank >= world) return nullptr;
  BnetTMesh* m = new BnetTMesh();
  m->rank = rank;
  m->world = world;
  m->dev = net_dev;
  m->listen = listen;
  return m;
}

const char* tmesh_error(BnetTMesh* m) { return m ? m->err : "null mesh"; }

// Every sender names itself on its connection: one 4-byte host message each way, before anything else.
static int identify_peers(BnetTMesh* m, uint64_t t0, int timeout_ms) {
  const int nc = m->world - 1;
  std::vector<int32_t> got(nc, -1);
  int32_t me = m->rank;
  std::vector<Request*> rq(nc, nullptr), sq(nc, nullptr);
  std::vector<unsigned char> rfin(nc, 0), sfin(nc, 0), rposted(nc, 0), sposted(nc, 0);
  int left = 2 * nc;
  while (left) {
    for (int c = 0; c < nc; c++) {
      if (!rposted[c]) {
        int st = m->recv[c]->irecv(&got[c], sizeof(int32_t), 0, nullptr, &rq[c]);
        if (st) return fail(m, "irecv of the peer's rank failed: %s", status_str(st));
        if (rq[c]) rposted[c] = 1;
      }
      if (!sposted[c]) {
        int st = m->send[c]->isend(&me, sizeof(int32_t), 0, nullptr, &sq[c]);
        if (st) return fail(m, "isend of the own rank failed: %s", status_str(st));
        if (sq[c]) sposted[c] = 1;
      }
      for (int side = 0; side < 2; side++) {
        std::vector<unsigned char>& fin = side ? sfin : rfin;
        if (fin[c] || !(side ? sposted[c] : rposted[c])) continue;
        int done = 0;
        size_t sz = 0;
        int st = (side ? m->send[c] : m->recv[c])->test(side ? sq[c] : rq[c], &done, &sz);
        if (st) return fail(m, "rank exchange failed: %s", status_str(st));
        if (done) { fin[c] = 1; left--; }
      }
    }
    if (left && timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull) return fail(m, "rank exchange timed out");
  }
  std::vector<unsigned char> seen(m->world, 0);
  m->recv_peer.assign(nc, -1);
  for (int c = 0; c < nc; c++) {
    if (got[c] < 0 || got[c] >= m->world || got[c] == m->rank || seen[got[c]])
      return fail(m, "incoming connection %d names rank %d (world %d): not a mesh", c, got[c], m->world);
    seen[got[c]] = 1;
    m->recv_peer[c] = got[c];
  }
  m->identified = true;
  return 0;
}

int tmesh_connect(BnetTMesh* m, const void* handles, int timeout_ms) {
  if (!m || !handles) return -1;
  const int n = m->world;
  const uint64_t t0 = now_ns();
  int next = 1 + (int)m->send.size();                   // offset of the next peer to connect to
  while ((int)m->send.size() < n - 1 || (int)m->recv.size() < n - 1) {
    bool moved = false;
    if ((int)m->send.size() < n - 1) {
      const int peer = (m->rank + next) % n;
      Comm* c = nullptr;
      int st = Engine::get().connect(m->dev, (const char*)handles + (size_t)peer * NCCL_NET_HANDLE_MAXSIZE, &c);
      if (st) return fail(m, "connect to rank %d failed: %s", peer, status_str(st));
      if (c) { m->send.push_back(c); next++; moved = true; }
    }
    if ((int)m->recv.size() < n - 1) {
      Comm* c = nullptr;
      int st = Engine::get().accept(m->listen, &c, false);
      if (st) return fail(m, "accept failed: %s", status_str(st));
      if (c) { m->recv.push_back(c); moved = true; }
    }
    if (timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull)
      return fail(m, "mesh connection timed out (%zu/%d out, %zu/%d in)", m->send.size(), n - 1, m->recv.size(), n - 1);
    if (!moved) usleep(200);
  }
  if (!m->identified && identify_peers(m, t0, timeout_ms)) return -1;
  if (!m->jflags) {
    void* dp = nullptr;
    m->jflags = (volatile uint64_t*)cuda::host_alloc_mapped(sizeof(uint64_t) * cuda::kMaxChunksPerJob, &dp);
    m->jflags_dev = (uint64_t*)dp;
    if (!m->jflags) return fail(m, "no memory for the executor completion words");
  }
  return 0;
}

MeshMr* tmesh_reg(BnetTMesh* m, void* data, size_t bytes, int ptr_type) {
  if (!m || !m->identified) return nullptr;
  MeshMr* mr = new MeshMr();
  mr->base = (char*)data;
  mr->bytes = bytes;
  mr->type = ptr_type;
  for (Comm* c : m->send) {
    MemHandle* h = nullptr;
    int st = c->reg_mr(data, bytes, ptr_type, &h);
    if (st) { fail(m, "regMr (send side) failed: %s", status_str(st)); tmesh_dereg(m, mr); return nullptr; }
    mr->smh.push_back(h);
  }
  for (Comm* c : m->recv) {
    MemHandle* h = nullptr;
    int st = c->reg_mr(data, bytes, ptr_type, &h);
    if (st) { fail(m, "regMr (receive side) failed: %s", status_str(st)); tmesh_dereg(m, mr); return nullptr; }
    mr->rmh.push_back(h);
  }
  return mr;
}

void tmesh_dereg(BnetTMesh* m, MeshMr* mr) {
  if (!m || !mr) return;
  for (size_t i = 0; i < mr->smh.size(); i++) m->send[i]->dereg_mr(mr->smh[i]);
  for (size_t i = 0; i < mr->rmh.size(); i++) m->recv[i]->dereg_mr(mr->rmh[i]);
  delete mr;
}

bool tmesh_mr_covers(const MeshMr* mr, const void* p, size_t bytes) {
  return mr && (const char*)p >= mr->base && (const char*)p + bytes <= mr->base + mr->bytes;
}

MeshOp* tmesh_op_start(BnetTMesh* m, int algo, const void* in, MeshMr* in_mr, void* out, MeshMr* out_mr, size_t count,
                       int in_dtype, int out_dtype, size_t piece_bytes, int inflight, int timeout_ms) {
  if (!m || !m->identified || !in_mr || !out_mr) return nullptr;
  if (m->active) { fail(m, "an all-reduce is still in flight on this mesh"); return nullptr; }
  if (algo != MESH_ONE_SHOT && algo != MESH_TWO_SHOT) { fail(m, "unknown algorithm %d", algo); return nullptr; }
  uint32_t add_op, init_op;
  if (in_dtype == 0 && out_dtype == 0) { add_op = cuda::OP_RED_ADD_F32; init_op = cuda::OP_COPY; }
  else if (in_dtype == 1 && out_dtype == 1) { add_op = cuda::OP_RED_ADD_BF16; init_op = cuda::OP_COPY; }
  else if (in_dtype == 1 && out_dtype == 0) { add_op = cuda::OP_ACC_BF16_TO_F32; init_op = cuda::OP_CAST_BF16_TO_F32; }
  else { fail(m, "unsupported dtype pair (%d -> %d)", in_dtype, out_dtype); return nullptr; }
  const size_t ies = in_dtype == 0 ? 4 : 2, oes = out_dtype == 0 ? 4 : 2;
  char* src = (char*)const_cast<void*>(in);
  char* dst = (char*)out;
  if (!tmesh_mr_covers(in_mr, src, count * ies)) { fail(m, "input outside the registered range"); return nullptr; }
  if (!tmesh_mr_covers(out_mr, dst, count * oes)) { fail(m, "output outside the registered range"); return nullptr; }
  // (16-byte alignment of both buffers puts every piece on the vector path of the executor ops; anything element-aligned
  //  still works, element by element)
  if (((uintptr_t)src % ies) || ((uintptr_t)dst % oes)) { fail(m, "buffers must be aligned to their element size"); return nullptr; }
  const bool in_place = src == dst;
  if (in_place && (algo != MESH_TWO_SHOT || ies != oes)) { fail(m, "in-place needs the two-shot algorithm and equal types"); return nullptr; }
  if (!in_place && src < dst + count * oes && dst < src + count * ies) { fail(m, "input and output overlap"); return nullptr; }
  int cuda_dev = 0;
  const bool host = in_mr->type == NCCL_PTR_HOST && out_mr->type == NCCL_PTR_HOST;
  if (in_mr->type != out_mr->type) { fail(m, "input and output must both be host memory or both be device memory"); return nullptr; }
  if (host && algo != MESH_TWO_SHOT) { fail(m, "host memory is reduced with the two-shot algorithm only"); return nullptr; }
  if (!host && !cuda::fake() && (!cuda::available() || !cuda::pointer_is_device(out, &cuda_dev))) {
    fail(m, "buffers registered as device memory must be device memory");
    return nullptr;
  }
  const int n = m->world, nc = n - 1;
  size_t piece = piece_bytes / ies / 64 * 64;             // elements per message; keeps both sides vector aligned
  if (piece < 1024) piece = 1024;
  if (inflight < 1) inflight = 1;
  if (inflight > kMaxRequests / 2) inflight = kMaxRequests / 2;

  MeshOp* op = new MeshOp();
  op->m = m;
  op->cuda_dev = cuda_dev;
  op->inflight = inflight;
  op->timeout_ms = timeout_ms;
  op->t0 = now_ns();
  op->rs.resize(nc);
  op->ss.resize(nc);
  op->host = host;
  op->in_dtype = in_dtype;
  op->out_dtype = out_dtype;
  // pieces of the element range [e0, e1) appended to a side's message list
  auto cut = [&](MeshSide& side, char* base, size_t es, size_t e0, size_t e1, uint32_t xop, MemHandle* mh, uint8_t phase) {
    for (size_t a = e0; a < e1; a += piece) {
      const size_t b = a + piece < e1 ? a + piece : e1;
      side.msgs.push_back(MeshMsg{base + a * es, (b - a) * es, xop, mh, phase});
    }
  };
  size_t seg = (count + n - 1) / n;
  seg = (seg + 63) / 64 * 64;                             // slice length in elements (the last slices may be short or empty)
  auto slice = [&](int s, size_t* e0, size_t* e1) {
    *e0 = (size_t)s * seg < count ? (size_t)s * seg : count;
    *e1 = (size_t)(s + 1) * seg < count ? (size_t)(s + 1) * seg : count;
  };
  size_t i0 = 0, i1 = count;                              // the local pass covers these elements
  if (algo == MESH_ONE_SHOT) {
    for (int c = 0; c < nc; c++) {
      cut(op->rs[c], dst, oes, 0, count, 0, out_mr->rmh[c], 0);
      cut(op->ss[c], src, ies, 0, count, add_op, in_mr->smh[c], 0);
    }
  } else {
    size_t m0, m1;
    slice(m->rank, &m0, &m1);
    i0 = m0;
    i1 = m1;
    if (host) {
      m->staging.resize(nc);
      for (auto& v : m->staging)
        if (v.size() < (m1 - m0) * ies + 64) v.resize((m1 - m0) * ies + 64);
    }
    for (int c = 0; c < nc; c++) {
      size_t q0, q1;
      if (!host) {
        cut(op->rs[c], dst, oes, m0, m1, 0, out_mr->rmh[c], 0);             // A: contributions to my slice (added by the sender)
      } else {
        // A, host mode: the peer's share of my slice lands in this connection's staging area (plain bytes of the input
        // type), piece by piece, and is added into out[slice] when it has arrived
        for (size_t a = m0; a < m1; a += piece) {
          const size_t b = a + piece < m1 ? a + piece : m1;
          MeshMsg g{m->staging[c].data() + (a - m0) * ies, (b - a) * ies, 0, nullptr, 0};
          g.acc = dst + a * oes;
          op->rs[c].msgs.push_back(g);
        }
      }
      slice(m->recv_peer[c], &q0, &q1);
      cut(op->rs[c], dst, oes, q0, q1, 0, out_mr->rmh[c], 1);               // B: that peer's finished slice
      slice((m->rank + 1 + c) % n, &q0, &q1);
      cut(op->ss[c], src, ies, q0, q1, host ? 0 : add_op, in_mr->smh[c], 0);   // A: my contribution to the peer's slice
      cut(op->ss[c], dst, oes, m0, m1, 0, out_mr->smh[c], 1);               // B: my finished slice
    }
  }
  for (int c = 0; c < nc; c++) {
    for (MeshSide* s : {&op->rs[c], &op->ss[c]}) {
      s->req.assign(s->msgs.size(), nullptr);
      s->fin.assign(s->msgs.size(), 0);
      op->remaining += s->msgs.size();
    }
    for (const MeshMsg& g : op->rs[c].msgs)
      if (g.phase == 0) op->a_recv_left++;
  }
  bool threads_serve_sends = false;
  for (int c = 0; c < nc; c++)
    threads_serve_sends = threads_serve_sends || strcmp(m->send[c]->transport(), "nvl") != 0 || strcmp(m->recv[c]->transport(), "nvl") != 0;
  if (in_place && threads_serve_sends) {
    op->a_send_left.assign(nc, 0);
    for (int c = 0; c < nc; c++)
      for (const MeshMsg& g : op->ss[c].msgs)
        if (g.phase == 0) op->a_send_left[c]++;
  }
  m->last_msgs = 0;
  m->last_bytes_sent = 0;
  op->out_bytes = (uint64_t)count * oes;
  op->span = Telemetry::get().span_begin(SPAN_COLL, (uint64_t)m->rank, ++m->op_seq, op->out_bytes);
  // ---- the local pass: out <- in over [i0, i1), finished before any receive is posted (nobody adds into garbage)
  op->stage = 1;
  if (host && !in_place && i1 > i0) {
    if (ies == oes) {
      memcpy(dst + i0 * oes, src + i0 * ies, (i1 - i0) * ies);
    } else {
      memset(dst + i0 * oes, 0, (i1 - i0) * oes);
      host_accumulate(dst + i0 * oes, src + i0 * ies, (i1 - i0) * ies, in_dtype, out_dtype);
    }
  } else if (!in_place && i1 > i0) {
    op->init_value = ++m->jseq;
    if (cuda::exec_transfer(cuda_dev, init_op, 1.0f, src + i0 * ies, dst + i0 * oes, (i1 - i0) * ies, m->jflags, m->jflags_dev,
                            op->init_value, &op->init_chunks) != 0) {
      fail(m, "the local initialisation pass failed");
      tmesh_op_free(op);
      return nullptr;
    }
    op->stage = 0;
  }
  m->active = op;
  return op;
}

int tmesh_op_step(MeshOp* op) {
  if (!op) return -1;
  if (op->stage == 2) return 1;
  if (op->stage < 0) return -1;
  BnetTMesh* m = op->m;
  const bool expired = op->timeout_ms > 0 && now_ns() - op->t0 > (uint64_t)op->timeout_ms * 1000000ull;
  if (op->stage == 0) {
    bool done = true;
    for (int c = 0; c < op->init_chunks; c++) done = done && m->jflags[c] == op->init_value;
    if (!done) {
      if (!cuda::fake()) cuda::exec_kick(op->cuda_dev);
      if (expired) { op->stage = -1; return fail(m, "the local initialisation pass timed out"); }
      return 0;
    }
    op->stage = 1;
  }
  bool moved = false;
  const int nc = m->world - 1;
  for (int c = 0; c < nc; c++) {
    MeshSide& R = op->rs[c];
    while (R.posted < R.msgs.size() && R.posted - R.done < (size_t)op->inflight) {
      const MeshMsg& g = R.msgs[R.posted];
      // in place, this peer's finished slice lands where my contribution to it was read: see MeshOp::a_send_left
      if (g.phase == 1 && !op->a_send_left.empty() && op->a_send_left[(m->recv_peer[c] - m->rank - 1 + m->world) % m->world]) break;
      Request* q = nullptr;
      int st = m->recv[c]->irecv(g.ptr, g.len, 0, g.mh, &q);
      if (st) { op->stage = -1; return fail(m, "irecv failed: %s", status_str(st)); }
      if (!q) break;
      R.req[R.posted++] = q;
      moved = true;
    }
    MeshSide& S = op->ss[c];
    while (S.posted < S.msgs.size() && S.posted - S.done < (size_t)op->inflight) {
      const MeshMsg& g = S.msgs[S.posted];
      if (g.phase == 1 && op->a_recv_left) break;         // my slice is not complete yet
      Request* q = nullptr;
      int st = g.op ? m->send[c]->isend_op(g.ptr, g.len, 0, g.mh, g.op, 1.0f, &q) : m->send[c]->isend(g.ptr, g.len, 0, g.mh, &q);
      if (st) { op->stage = -1; return fail(m, "isend failed: %s (the fused isend needs the NVLink transport)", status_str(st)); }
      if (!q) break;
      S.req[S.posted++] = q;
      m->last_msgs++;
      m->last_bytes_sent += g.len;
      moved = true;
    }
    for (int side = 0; side < 2; side++) {
      MeshSide& X = side ? op->ss[c] : op->rs[c];
      Comm* comm = side ? m->send[c] : m->recv[c];
      for (size_t i = X.done; i < X.posted; i++) {
        if (X.fin[i]) continue;
        int done = 0;
        size_t sz = 0;
        int st = comm->test(X.req[i], &done, &sz);
        if (st) { op->stage = -1; return fail(m, "%s %zu failed: %s", side ? "send" : "receive", i, status_str(st)); }
        if (done) {
          X.fin[i] = 1;
          op->remaining--;
          if (side && X.msgs[i].phase == 0 && !op->a_send_left.empty()) op->a_send_left[c]--;
          if (!side && X.msgs[i].phase == 0) {
            if (X.msgs[i].acc) host_accumulate(X.msgs[i].acc, X.msgs[i].ptr, X.msgs[i].len, op->in_dtype, op->out_dtype);
            op->a_recv_left--;
          }
          moved = true;
        }
      }
      while (X.done < X.posted && X.fin[X.done]) X.done++;
    }
  }
  op->idle = moved ? 0 : op->idle + 1;
  if (op->remaining == 0) {
    op->stage = 2;
    if (op->span) Telemetry::get().span_end(op->span, op->out_bytes);
    op->span = 0;
    return 1;
  }
  if (!moved && expired) {
    op->stage = -1;
    return fail(m, "all-reduce over the mesh timed out with %zu request(s) outstanding", op->remaining);
  }
  return 0;
}

void tmesh_op_free(MeshOp* op) {
  if (!op) return;
  if (op->span) Telemetry::get().span_end(op->span, 0);       // (failed or abandoned: closed with 0 bytes)
  if (op->m && op->m->active == op) op->m->active = nullptr;
  delete op;
}

void tmesh_destroy(BnetTMesh* m) {
  if (!m) return;
  if (m->active) tmesh_op_free(m->active);
  tmesh_dereg(m, m->in_mr);
  if (m->out_mr != m->in_mr) tmesh_dereg(m, m->out_mr);
  // (the completion words are deliberately not freed: cudaFreeHost waits for running kernels)
  for (Comm* c : m->send) delete c;
  for (Comm* c : m->recv) delete c;
  delete m->listen;
  delete m;
}

}  // namespace bnet

// ---------------------------------------------------------------------------------------------------------------------------
// C API (bagua_net_b200/parallel/transport_ring.py: MeshCore / TransportMesh)
BNET_API const char* bnet_tmesh_last_error(BnetTMesh* m) { return tmesh_error(m); }

// Step 1 on every rank: listen.  `handle_out` (>= 128 bytes) goes to EVERY other rank (out of band).
BNET_API int bnet_tmesh_create(int rank, int world, int net_dev, void* handle_out, BnetTMesh** out) {
  if (!out || !handle_out || world < 2 || world > 64 || rank < 0 || rank >= world) return -1;
  if (Engine::get().init()) return -1;
  if (net_dev < 0 || net_dev >= Engine::get().ndev()) return -1;
  ListenComm* l = nullptr;
  if (Engine::get().listen(net_dev, handle_out, NCCL_NET_HANDLE_MAXSIZE, &l)) return -1;
  BnetTMesh* m = tmesh_new(l, rank, world, net_dev);
  if (!m) { delete l; return -1; }
  *out = m;
  return 0;
}

// Step 2: `handles` = world x 128 bytes, entry r produced by rank r.  Connects to every peer and accepts every peer.
BNET_API int bnet_tmesh_connect(BnetTMesh* m, const void* handles, int timeout_ms) { return tmesh_connect(m, handles, timeout_ms); }

BNET_API const char* bnet_tmesh_transport(BnetTMesh* m) { return m && !m->send.empty() ? m->send[0]->transport() : ""; }

// Step 3: the input (read by this rank's kernels) and the output (accumulated into by every peer's kernels, and read by
// this rank's kernels in the all-gather phase).  in == out registers one buffer for in-place two-shot all-reduces.
// ptr_type: NCCL_PTR_CUDA (2: device memory, NVLink transport, reduction fused into the isends) or NCCL_PTR_HOST (1: host
// memory over ANY transport — TCP between hosts included; two-shot only, the reduce-scatter pieces are added on the host).
BNET_API int bnet_tmesh_register2(BnetTMesh* m, void* in, size_t in_bytes, void* out, size_t out_bytes, int ptr_type) {
  if (!m || !m->identified || m->active || (ptr_type != NCCL_PTR_CUDA && ptr_type != NCCL_PTR_HOST)) return -1;
  tmesh_dereg(m, m->in_mr);
  if (m->out_mr != m->in_mr) tmesh_dereg(m, m->out_mr);
  m->in_mr = m->out_mr = nullptr;
  m->in_mr = tmesh_reg(m, in, in_bytes, ptr_type);
  if (!m->in_mr) return -1;
  if (in == out && in_bytes == out_bytes) {
    m->out_mr = m->in_mr;
    return 0;
  }
  m->out_mr = tmesh_reg(m, out, out_bytes, ptr_type);
  return m->out_mr ? 0 : -1;
}
BNET_API int bnet_tmesh_register(BnetTMesh* m, void* in, size_t in_bytes, void* out, size_t out_bytes) {
  return bnet_tmesh_register2(m, in, in_bytes, out, out_bytes, NCCL_PTR_CUDA);
}

// out = sum over ranks of in.  in_dtype / out_dtype: 0 = fp32, 1 = bf16; (fp32, fp32), (bf16, bf16) and (bf16 in, fp32 out:
// the accumulate-while-widening op) are supported.  algo: 0 = one-shot, 1 = two-shot (see the top of this file).  Every rank
// calls with the same arguments after making sure `in` is complete (stream synchronised).
BNET_API int bnet_tmesh_allreduce2(BnetTMesh* m, const void* in, void* out, size_t count, int in_dtype, int out_dtype, int algo,
                                   size_t piece_bytes, int inflight, int timeout_ms) {
  if (!m || !m->in_mr || !m->out_mr) return -1;
  MeshOp* op = tmesh_op_start(m, algo, in, m->in_mr, out, m->out_mr, count, in_dtype, out_dtype, piece_bytes, inflight, timeout_ms);
  if (!op) return -1;
  int st;
  IdleBackoff backoff;
  while ((st = tmesh_op_step(op)) == 0) {
    // Poll hot while things move (a GPU kernel's completion word arrives within microseconds); when nothing has moved for
    // a while the peers' threads — TCP stream workers, other ranks of an oversubscribed host — need the core more than we do
    // (coll/backoff.h)
    backoff.step(op->idle == 0);
  }
  tmesh_op_free(op);
  return st == 1 ? 0 : -1;
}

BNET_API int bnet_tmesh_allreduce(BnetTMesh* m, const void* in, void* out, size_t count, int in_dtype, int out_dtype,
                                  size_t piece_bytes, int inflight, int timeout_ms) {
  return bnet_tmesh_allreduce2(m, in, out, count, in_dtype, out_dtype, MESH_ONE_SHOT, piece_bytes, inflight, timeout_ms);
}

BNET_API void bnet_tmesh_stats(BnetTMesh* m, unsigned long long* msgs, unsigned long long* bytes_sent) {
  if (msgs) *msgs = m ? m->last_msgs : 0;
  if (bytes_sent) *bytes_sent = m ? m->last_bytes_sent : 0;
}

BNET_API void bnet_tmesh_destroy(BnetTMesh* m) { tmesh_destroy(m); }
