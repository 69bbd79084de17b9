// Ring all-reduce that RIDES THE TRANSPORT: every hop is an isend/irecv pair on the plugin's own connections, and the
// reduction is done by the sending kernel while it moves the data ("fused isend", K4 of SURVEY.md section 2.6) — the
// receiver's posted buffer region is accumulated into (red.global.add over NVLink), so no elementwise reduce kernel
// exists anywhere and nothing is staged: the user buffer itself is the communication buffer.
//
// The reference moves bytes for NCCL and leaves the reduction to NCCL's kernels (reference README.md:88, SURVEY.md
// section 2.5 "collective algorithms: absent"); this is the collective the north star asks to ride the new transport.
//
//   rank r, world n, buffer cut into n segments, every segment into pieces of <= piece_bytes:
//     reduce-scatter step s = 0..n-2 : send segment (r - s)     with OP_RED_ADD_*  -> accumulated into next's copy
//                                      recv segment (r - s - 1)
//     all-gather    step s = 0..n-2 : send segment (r + 1 - s)  with OP_COPY
//                                      recv segment (r - s)
//   piece p of step g may leave once piece p of step g-1 has ARRIVED (it is the same memory): pieces pipeline around the
//   ring with up to `inflight` requests per direction in flight, striped over the executor's clusters by the transport.
//
// Host-driven (one polling thread per rank, like NCCL's proxy); every request goes through Comm::isend_op / irecv / test.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "core/engine.h"
#include "cuda/exec_ops.h"

using namespace bnet;

#define BNET_API extern "C" __attribute__((visibility("default")))

struct BnetTRing {
  int rank = 0, world = 1, dev = 0;
  ListenComm* listen = nullptr;
  Comm* send_next = nullptr;
  Comm* recv_prev = nullptr;
  MemHandle* mh_send = nullptr;
  MemHandle* mh_recv = nullptr;
  char* reg_base = nullptr;
  size_t reg_bytes = 0;
  char err[256] = {0};
  // statistics of the last all-reduce
  uint64_t last_msgs = 0, last_bytes_sent = 0;
};

namespace {
int fail(BnetTRing* r, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(BnetTRing* r, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(r->err, sizeof(r->err), fmt, ap);
  va_end(ap);
  BNET_WARN("transport ring: %s", r->err);
  return -1;
}
}  // namespace

BNET_API const char* bnet_tring_last_error(BnetTRing* r) { return r ? r->err : "null ring"; }

// Step 1 on every rank: listen.  `handle_out` (>= 128 bytes) goes to the PREVIOUS rank in the ring (out of band).
BNET_API int bnet_tring_create(int rank, int world, int net_dev, void* handle_out, BnetTRing** out) {
  if (!out || !handle_out || world < 2 || rank < 0 || rank >= world) return -1;
  int st = Engine::get().init();
  if (st) return -1;
  BnetTRing* r = new BnetTRing();
  r->rank = rank;
  r->world = world;
  r->dev = net_dev;
  if (net_dev < 0 || net_dev >= Engine::get().ndev()) { delete r; return -1; }
  st = Engine::get().listen(net_dev, handle_out, NCCL_NET_HANDLE_MAXSIZE, &r->listen);
  if (st) { delete r; return -1; }
  *out = r;
  return 0;
}

// Step 2: `next_handle` is the handle the NEXT rank produced.  Connects to it and accepts the previous rank.
BNET_API int bnet_tring_connect(BnetTRing* r, const void* next_handle, int timeout_ms) {
  if (!r || !next_handle) return -1;
  const uint64_t t0 = now_ns();
  while (!r->send_next || !r->recv_prev) {
    if (!r->send_next) {
      int st = Engine::get().connect(r->dev, next_handle, &r->send_next);
      if (st) return fail(r, "connect to the next rank failed: %s", status_str(st));
    }
    if (!r->recv_prev) {
      int st = Engine::get().accept(r->listen, &r->recv_prev, false);
      if (st) return fail(r, "accept from the previous rank failed: %s", status_str(st));
    }
    if (timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull) return fail(r, "ring connection timed out");
    if (!r->send_next || !r->recv_prev) usleep(200);
  }
  return 0;
}

BNET_API const char* bnet_tring_transport(BnetTRing* r) { return r && r->send_next ? r->send_next->transport() : ""; }

// Step 3: the buffer all-reduces will run on (device memory; regMr exports it to the previous rank's kernels).
BNET_API int bnet_tring_register(BnetTRing* r, void* buf, size_t bytes) {
  if (!r || !r->send_next || !r->recv_prev) return -1;
  if (r->mh_send) { r->send_next->dereg_mr(r->mh_send); r->mh_send = nullptr; }
  if (r->mh_recv) { r->recv_prev->dereg_mr(r->mh_recv); r->mh_recv = nullptr; }
  int st = r->send_next->reg_mr(buf, bytes, NCCL_PTR_CUDA, &r->mh_send);
  if (st) return fail(r, "regMr (send side) failed: %s", status_str(st));
  st = r->recv_prev->reg_mr(buf, bytes, NCCL_PTR_CUDA, &r->mh_recv);
  if (st) return fail(r, "regMr (receive side) failed: %s", status_str(st));
  r->reg_base = (char*)buf;
  r->reg_bytes = bytes;
  return 0;
}

// In-place sum all-reduce of `count` elements at `buf` (inside the registered range).  dtype: 0 = fp32, 1 = bf16.
// Every rank calls it with the same arguments; the caller has made sure the buffer's contents are complete (stream
// synchronised) — the same contract NCCL gives a net plugin.  Returns 0, or -1 (bnet_tring_last_error).
BNET_API int bnet_tring_allreduce(BnetTRing* r, void* buf, size_t count, int dtype, size_t piece_bytes, int inflight,
                                  int timeout_ms) {
  if (!r || !r->mh_send || !r->mh_recv) return -1;
  const size_t es = dtype == 0 ? 4 : 2;
  const uint32_t red_op = dtype == 0 ? cuda::OP_RED_ADD_F32 : cuda::OP_RED_ADD_BF16;
  char* base = (char*)buf;
  if (base < r->reg_base || base + count * es > r->reg_base + r->reg_bytes) return fail(r, "buffer outside the registered range");
  if (((uintptr_t)base & 15) != 0) return fail(r, "buffer must be 16-byte aligned");
  const int n = r->world;
  if (piece_bytes < 4096) piece_bytes = 4096;
  piece_bytes = piece_bytes / 64 * 64;
  if (inflight < 1) inflight = 1;
  if (inflight > kMaxRequests / 2) inflight = kMaxRequests / 2;
  // segments: equal, 64-byte aligned (vector-aligned on both sides for every op); the last one may be short or empty
  const size_t total = count * es;
  size_t seg = (total + n - 1) / n;
  seg = (seg + 63) / 64 * 64;
  const size_t np = (seg + piece_bytes - 1) / piece_bytes;     // pieces per segment (same on every rank)
  auto piece = [&](int segment, size_t p, size_t* off, size_t* len) {
    size_t s0 = (size_t)segment * seg, s1 = s0 + seg;
    if (s0 > total) s0 = total;
    if (s1 > total) s1 = total;
    size_t a = s0 + p * piece_bytes, b = a + piece_bytes;
    if (a > s1) a = s1;
    if (b > s1) b = s1;
    *off = a;
    *len = b - a;
  };
  const int steps = 2 * (n - 1);
  const size_t M = (size_t)steps * np;
  auto send_seg = [&](int g) { return g < n - 1 ? ((r->rank - g) % n + n) % n : ((r->rank + 1 - (g - (n - 1))) % n + n) % n; };
  auto recv_seg = [&](int g) { return g < n - 1 ? ((r->rank - g - 1) % n + n) % n : ((r->rank - (g - (n - 1))) % n + n) % n; };
  std::vector<Request*> rreq(M, nullptr), sreq(M, nullptr);
  std::vector<unsigned char> rdone(M, 0), sdone(M, 0);
  size_t posted_r = 0, posted_s = 0, done_r = 0, done_s = 0, head_r = 0, head_s = 0;
  r->last_msgs = 0;
  r->last_bytes_sent = 0;
  const uint64_t t0 = now_ns();
  while (done_r < M || done_s < M) {
    bool moved = false;
    // receives: posted in message order (the transport matches strictly FIFO per connection)
    while (posted_r < M && posted_r - done_r < (size_t)inflight) {
      const int g = (int)(posted_r / np);
      size_t off, len;
      piece(recv_seg(g), posted_r % np, &off, &len);
      Request* q = nullptr;
      int st = r->recv_prev->irecv(base + off, len, 0, r->mh_recv, &q);
      if (st) return fail(r, "irecv failed: %s", status_str(st));
      if (!q) break;
      rreq[posted_r++] = q;
      moved = true;
    }
    // sends: piece p of step g leaves once piece p of step g-1 has arrived (same memory)
    while (posted_s < M && posted_s - done_s < (size_t)inflight) {
      const int g = (int)(posted_s / np);
      const size_t p = posted_s % np;
      if (g > 0 && !rdone[(size_t)(g - 1) * np + p]) break;
      size_t off, len;
      piece(send_seg(g), p, &off, &len);
      Request* q = nullptr;
      const uint32_t op = (g < n - 1 && len > 0) ? red_op : (uint32_t)cuda::OP_COPY;
      int st = r->send_next->isend_op(base + off, len, 0, r->mh_send, op, 1.0f, &q);
      if (st) return fail(r, "isend_op failed: %s", status_str(st));
      if (!q) break;
      sreq[posted_s++] = q;
      r->last_msgs++;
      r->last_bytes_sent += len;
      moved = true;
    }
    // completions (out of order is fine; heads advance over finished ones)
    for (size_t i = head_r; i < posted_r; i++) {
      if (rdone[i]) continue;
      int done = 0;
      size_t sz = 0;
      int st = r->recv_prev->test(rreq[i], &done, &sz);
      if (st) return fail(r, "receive %zu failed: %s", i, status_str(st));
      if (done) { rdone[i] = 1; done_r++; moved = true; }
    }
    while (head_r < posted_r && rdone[head_r]) head_r++;
    for (size_t i = head_s; i < posted_s; i++) {
      if (sdone[i]) continue;
      int done = 0;
      size_t sz = 0;
      int st = r->send_next->test(sreq[i], &done, &sz);
      if (st) return fail(r, "send %zu failed: %s", i, status_str(st));
      if (done) { sdone[i] = 1; done_s++; moved = true; }
    }
    while (head_s < posted_s && sdone[head_s]) head_s++;
    if (!moved && timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull)
      return fail(r, "all-reduce timed out: %zu/%zu receives, %zu/%zu sends done", done_r, M, done_s, M);
  }
  return 0;
}

BNET_API void bnet_tring_stats(BnetTRing* r, unsigned long long* msgs, unsigned long long* bytes_sent) {
  if (msgs) *msgs = r ? r->last_msgs : 0;
  if (bytes_sent) *bytes_sent = r ? r->last_bytes_sent : 0;
}

BNET_API void bnet_tring_destroy(BnetTRing* r) {
  if (!r) return;
  if (r->mh_send && r->send_next) r->send_next->dereg_mr(r->mh_send);
  if (r->mh_recv && r->recv_prev) r->recv_prev->dereg_mr(r->mh_recv);
  delete r->send_next;
  delete r->recv_prev;
  delete r->listen;
  delete r;
}
