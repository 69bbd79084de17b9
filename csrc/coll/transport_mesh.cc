// One-shot all-reduce that RIDES THE TRANSPORT: a full mesh of plugin connections, every rank sends its whole input to every
// peer ONCE, and the sending kernel accumulates it into the peer's output buffer while it moves the data (fused isend,
// K4 / K4+K5 of SURVEY.md section 2.6).  One network step instead of the ring's 2(n-1): the latency-optimal shape for
// messages up to a few MiB, where the ring (csrc/coll/transport_ring.cc) pays ~30 us per step; the ring wins once
// (n-1) x size / link bandwidth exceeds that.
//
//   out <- in                      local executor pass (copy, or bf16 -> fp32 cast when the output is wider)
//   irecv(out) on every incoming connection            (posted only after the local pass: nobody adds into garbage)
//   isend_op(in, OP_RED_ADD_F32 | OP_RED_ADD_BF16 | OP_ACC_BF16_TO_F32) on every outgoing connection, cut into pieces,
//            `inflight` requests per connection — n-1 kernels accumulate into each output concurrently (red.global.add)
//
// The sum order is not fixed (n-1 senders race), so results are bit-identical across ranks only when the additions are
// exact (integers, or fp32 accumulation of bf16 inputs that does not round) — use the ring for reproducible floats.
// The reference has no collectives (reference README.md:88); this and the ring are the collectives
// section 7.2 step 4 of the survey asks to build over the transport's isend / irecv.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "core/engine.h"
#include "cuda/cuda_iface.h"
#include "cuda/exec_ops.h"
#include "cuda/nvl_exec.h"

using namespace bnet;

#define BNET_API extern "C" __attribute__((visibility("default")))

struct BnetTMesh {
  int rank = 0, world = 1, dev = 0;
  ListenComm* listen = nullptr;
  std::vector<Comm*> send;          // world - 1 outgoing connections (order: rank+1, rank+2, ...: spreads the first hits)
  std::vector<Comm*> recv;          // world - 1 incoming connections (arrival order; all are treated alike)
  std::vector<MemHandle*> mh_in, mh_out;
  char* in_base = nullptr;
  size_t in_bytes = 0;
  char* out_base = nullptr;
  size_t out_bytes = 0;
  volatile uint64_t* jflags = nullptr;
  uint64_t* jflags_dev = nullptr;
  uint64_t jseq = 0;
  char err[256] = {0};
  uint64_t last_msgs = 0, last_bytes_sent = 0;
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
}  // namespace

BNET_API const char* bnet_tmesh_last_error(BnetTMesh* m) { return m ? m->err : "null mesh"; }

// Step 1 on every rank: listen.  `handle_out` (>= 128 bytes) goes to EVERY other rank (out of band).
BNET_API int bnet_tmesh_create(int rank, int world, int net_dev, void* handle_out, BnetTMesh** out) {
  if (!out || !handle_out || world < 2 || world > 64 || rank < 0 || rank >= world) return -1;
  if (Engine::get().init()) return -1;
  if (net_dev < 0 || net_dev >= Engine::get().ndev()) return -1;
  BnetTMesh* m = new BnetTMesh();
  m->rank = rank;
  m->world = world;
  m->dev = net_dev;
  if (Engine::get().listen(net_dev, handle_out, NCCL_NET_HANDLE_MAXSIZE, &m->listen)) { delete m; return -1; }
  *out = m;
  return 0;
}

// Step 2: `handles` = world x 128 bytes, entry r produced by rank r.  Connects to every peer and accepts every peer.
BNET_API int bnet_tmesh_connect(BnetTMesh* m, const void* handles, int timeout_ms) {
  if (!m || !handles) return -1;
  const int n = m->world;
  const uint64_t t0 = now_ns();
  int next = 1;                       // offset of the next peer to connect to
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
  return 0;
}

BNET_API const char* bnet_tmesh_transport(BnetTMesh* m) { return m && !m->send.empty() ? m->send[0]->transport() : ""; }

// Step 3: the input (read by this rank's kernels) and the output (accumulated into by every peer's kernels).
BNET_API int bnet_tmesh_register(BnetTMesh* m, void* in, size_t in_bytes, void* out, size_t out_bytes) {
  if (!m || (int)m->send.size() != m->world - 1 || (int)m->recv.size() != m->world - 1) return -1;
  for (size_t i = 0; i < m->mh_in.size(); i++) m->send[i]->dereg_mr(m->mh_in[i]);
  for (size_t i = 0; i < m->mh_out.size(); i++) m->recv[i]->dereg_mr(m->mh_out[i]);
  m->mh_in.clear();
  m->mh_out.clear();
  for (Comm* c : m->send) {
    MemHandle* h = nullptr;
    int st = c->reg_mr(in, in_bytes, NCCL_PTR_CUDA, &h);
    if (st) return fail(m, "regMr (input) failed: %s", status_str(st));
    m->mh_in.push_back(h);
  }
  for (Comm* c : m->recv) {
    MemHandle* h = nullptr;
    int st = c->reg_mr(out, out_bytes, NCCL_PTR_CUDA, &h);
    if (st) return fail(m, "regMr (output) failed: %s", status_str(st));
    m->mh_out.push_back(h);
  }
  m->in_base = (char*)in;
  m->in_bytes = in_bytes;
  m->out_base = (char*)out;
  m->out_bytes = out_bytes;
  if (!m->jflags) {
    void* dp = nullptr;
    m->jflags = (volatile uint64_t*)cuda::host_alloc_mapped(sizeof(uint64_t) * cuda::kMaxChunksPerJob, &dp);
    m->jflags_dev = (uint64_t*)dp;
    if (!m->jflags) return fail(m, "no memory for the executor completion words");
  }
  return 0;
}

// out = sum over ranks of in.  in_dtype / out_dtype: 0 = fp32, 1 = bf16; (fp32, fp32), (bf16, bf16) and (bf16 in, fp32 out:
// the accumulate-while-widening op) are supported.  `in` and `out` are different buffers inside the registered ranges; every
// rank calls with the same arguments after making sure `in` is complete (stream synchronised).
BNET_API int bnet_tmesh_allreduce(BnetTMesh* m, const void* in, void* out, size_t count, int in_dtype, int out_dtype,
                                  size_t piece_bytes, int inflight, int timeout_ms) {
  if (!m || m->mh_in.empty() || m->mh_out.empty()) return -1;
  uint32_t add_op, init_op;
  if (in_dtype == 0 && out_dtype == 0) { add_op = cuda::OP_RED_ADD_F32; init_op = cuda::OP_COPY; }
  else if (in_dtype == 1 && out_dtype == 1) { add_op = cuda::OP_RED_ADD_BF16; init_op = cuda::OP_COPY; }
  else if (in_dtype == 1 && out_dtype == 0) { add_op = cuda::OP_ACC_BF16_TO_F32; init_op = cuda::OP_CAST_BF16_TO_F32; }
  else return fail(m, "unsupported dtype pair (%d -> %d)", in_dtype, out_dtype);
  const size_t ies = in_dtype == 0 ? 4 : 2, oes = out_dtype == 0 ? 4 : 2;
  const char* src = (const char*)in;
  char* dst = (char*)out;
  if (src < m->in_base || src + count * ies > m->in_base + m->in_bytes) return fail(m, "input outside the registered range");
  if (dst < m->out_base || dst + count * oes > m->out_base + m->out_bytes) return fail(m, "output outside the registered range");
  if ((((uintptr_t)src) & 63) || (((uintptr_t)dst) & 63)) return fail(m, "buffers must be 64-byte aligned");
  if (src < dst + count * oes && dst < src + count * ies) return fail(m, "input and output overlap");
  int cuda_dev = 0;
  if (!cuda::fake() && (!cuda::available() || !cuda::pointer_is_device(out, &cuda_dev))) return fail(m, "buffers must be device memory");
  const int n = m->world, np_conn = n - 1;
  if (piece_bytes < 4096) piece_bytes = 4096;
  piece_bytes = piece_bytes / 256 * 256;                  // source bytes per message; keeps both sides vector aligned
  if (inflight < 1) inflight = 1;
  if (inflight > kMaxRequests / 2) inflight = kMaxRequests / 2;
  const size_t total = count * ies;
  const size_t np = total ? (total + piece_bytes - 1) / piece_bytes : 1;
  m->last_msgs = 0;
  m->last_bytes_sent = 0;
  const uint64_t t0 = now_ns();
  // ---- 1) out <- in (local pass), finished before any receive is posted
  {
    const uint64_t value = ++m->jseq;
    int nchunks = 0;
    if (total) {
      if (cuda::exec_transfer(cuda_dev, init_op, 1.0f, src, dst, total, m->jflags, m->jflags_dev, value, &nchunks) != 0)
        return fail(m, "the local initialisation pass failed");
      for (;;) {
        bool done = true;
        for (int c = 0; c < nchunks; c++) done = done && m->jflags[c] == value;
        if (done) break;
        if (!cuda::fake()) cuda::exec_kick(cuda_dev);
        if (timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull) return fail(m, "the local initialisation pass timed out");
      }
    }
  }
  // ---- 2) every connection carries np messages; receives and sends are posted in message order per connection
  struct Side {
    size_t posted = 0, done = 0;
    std::vector<Request*> req;
    std::vector<unsigned char> fin;
  };
  std::vector<Side> rs(np_conn), ss(np_conn);
  for (int c = 0; c < np_conn; c++) {
    rs[c].req.assign(np, nullptr); rs[c].fin.assign(np, 0);
    ss[c].req.assign(np, nullptr); ss[c].fin.assign(np, 0);
  }
  auto piece = [&](size_t p, size_t* off, size_t* len) {
    size_t a = p * piece_bytes, b = a + piece_bytes;
    if (a > total) a = total;
    if (b > total) b = total;
    *off = a;
    *len = b - a;
  };
  size_t remaining = 2 * (size_t)np_conn * np;
  while (remaining) {
    bool moved = false;
    for (int c = 0; c < np_conn; c++) {
      Side& R = rs[c];
      while (R.posted < np && R.posted - R.done < (size_t)inflight) {
        size_t off, len;
        piece(R.posted, &off, &len);
        const size_t ooff = off / ies * oes, olen = len / ies * oes;
        Request* q = nullptr;
        int st = m->recv[c]->irecv(dst + ooff, olen, 0, m->mh_out[c], &q);
        if (st) return fail(m, "irecv failed: %s", status_str(st));
        if (!q) break;
        R.req[R.posted++] = q;
        moved = true;
      }
      Side& S = ss[c];
      while (S.posted < np && S.posted - S.done < (size_t)inflight) {
        size_t off, len;
        piece(S.posted, &off, &len);
        Request* q = nullptr;
        int st = len ? m->send[c]->isend_op(src + off, len, 0, m->mh_in[c], add_op, 1.0f, &q)
                     : m->send[c]->isend(src + off, 0, 0, m->mh_in[c], &q);
        if (st) return fail(m, "isend_op failed: %s (the fused isend needs the NVLink transport)", status_str(st));
        if (!q) break;
        S.req[S.posted++] = q;
        m->last_msgs++;
        m->last_bytes_sent += len;
        moved = true;
      }
      for (int side = 0; side < 2; side++) {
        Side& X = side ? ss[c] : rs[c];
        Comm* comm = side ? m->send[c] : m->recv[c];
        for (size_t i = X.done; i < X.posted; i++) {
          if (X.fin[i]) continue;
          int done = 0;
          size_t sz = 0;
          int st = comm->test(X.req[i], &done, &sz);
          if (st) return fail(m, "%s %zu failed: %s", side ? "send" : "receive", i, status_str(st));
          if (done) { X.fin[i] = 1; remaining--; moved = true; }
        }
        while (X.done < X.posted && X.fin[X.done]) X.done++;
      }
    }
    if (!moved && timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull)
      return fail(m, "one-shot all-reduce timed out with %zu request(s) outstanding", remaining);
  }
  return 0;
}

BNET_API void bnet_tmesh_stats(BnetTMesh* m, unsigned long long* msgs, unsigned long long* bytes_sent) {
  if (msgs) *msgs = m ? m->last_msgs : 0;
  if (bytes_sent) *bytes_sent = m ? m->last_bytes_sent : 0;
}

BNET_API void bnet_tmesh_destroy(BnetTMesh* m) {
  if (!m) return;
  for (size_t i = 0; i < m->mh_in.size(); i++) m->send[i]->dereg_mr(m->mh_in[i]);
  for (size_t i = 0; i < m->mh_out.size(); i++) m->recv[i]->dereg_mr(m->mh_out[i]);
  for (Comm* c : m->send) delete c;
  for (Comm* c : m->recv) delete c;
  delete m->listen;
  delete m;
}
