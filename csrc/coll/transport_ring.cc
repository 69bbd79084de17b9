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

#include "coll/backoff.h"
#include "core/engine.h"
#include "core/telemetry.h"
#include "cuda/cuda_iface.h"
#include "cuda/exec_ops.h"
#include "cuda/nvl_exec.h"

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
  // compressed all-reduce: the wire-format mirror of the data buffer (registered with both connections) and the completion
  // words of the local executor jobs (quantise / dequantise / accumulate passes)
  MemHandle* mh_wsend = nullptr;
  MemHandle* mh_wrecv = nullptr;
  char* wire_base = nullptr;
  size_t wire_bytes = 0;
  volatile uint64_t* jflags = nullptr;
  uint64_t* jflags_dev = nullptr;
  uint64_t jseq = 0;
  uint64_t op_seq = 0;             // all-reduces started on this ring (span ids)
};

namespace {
// In place, the receive of an all-gather piece writes the memory a reduce-scatter send read n-1 steps earlier.  The bytes can
// only arrive after that send's data went all the way round the ring, so posting the receive early is harmless — by causality
// through the network, which no tool (and no transport with its own idea of "send complete") can see.  Over transports whose
// sends are served by other threads (TCP) the receive is therefore posted only once that send has completed locally: an
// explicit happens-before, a poll iteration late at worst.  The NVLink path (kernels on both ends, validated on hardware as
// it is) keeps posting ahead.
bool orders_receives_after_own_sends(const BnetTRing* r) {
  return strcmp(r->send_next->transport(), "nvl") != 0 || strcmp(r->recv_prev->transport(), "nvl") != 0;
}

int fail(BnetTRing* r, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(BnetTRing* r, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(r->err, sizeof(r->err), fmt, ap);
  va_end(ap);
  BNET_WARN("transport ring: %s", r->err);
  return -1;
}

// "coll-<rank>" span around one all-reduce (the isend / irecv spans of its messages fall inside it); closed with the byte
// count on success, with 0 when the call leaves early
struct CollSpan {
  uint64_t id, bytes;
  bool ok = false;
  CollSpan(int rank, uint64_t seq, uint64_t nbytes) : id(Telemetry::get().span_begin(SPAN_COLL, (uint64_t)rank, seq, nbytes)), bytes(nbytes) {}
  ~CollSpan() { if (id) Telemetry::get().span_end(id, ok ? bytes : 0); }
};
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
  CollSpan span(r->rank, ++r->op_seq, total);
  IdleBackoff backoff;
  const bool send_after_own_send = orders_receives_after_own_sends(r);
  while (done_r < M || done_s < M) {
    bool moved = false;
    // receives: posted in message order (the transport matches strictly FIFO per connection)
    while (posted_r < M && posted_r - done_r < (size_t)inflight) {
      const int g = (int)(posted_r / np);
      // an all-gather receive lands where the reduce-scatter send of n-1 steps earlier read (see send_after_own_send)
      if (send_after_own_send && g >= n - 1 && !sdone[posted_r - (size_t)(n - 1) * np]) break;
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
    // poll hot while things move; when nothing has for a while, the transport's worker threads (or other ranks of an
    // oversubscribed host) need the core more than this loop does (coll/backoff.h)
    backoff.step(moved);
  }
  span.ok = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Compressed all-reduce (K5 of SURVEY.md section 2.6: "dtype casts / compression" fused into the transport).  fp32 data, a
// narrower format on the wire: bf16 (2x fewer bytes), fp8 e4m3 or e5m2 (4x), scaled and saturating.
//
//   `wire` is a buffer of count * wire_elsize bytes that mirrors the data buffer segment by segment (same element offsets):
//   compressed pieces are received into it, and — on transports without the fused isend — quantised into it before leaving.
//
//   reduce-scatter hop : fused   isend_op(OP_CAST_F32_TO_<w>, scale) straight from the fp32 segment into the NEXT rank's wire
//                                mirror — one kernel quantises and moves, the link carries the narrow format
//                        unfused quantise locally into the own mirror (executor pass), plain isend of the narrow bytes
//                                (any transport: TCP between hosts is where compression pays most)
//                        receiver: OP_ACC_<w>_TO_F32 (1/scale) from its mirror into its fp32 segment, then the piece may go on
//   all-gather         : the owner of a finished segment quantises it ONCE into its mirror and also dequantises it back in
//                        place, so that every rank ends with the same bits; the narrow piece then travels round the ring
//                        unchanged (plain isend from mirror to mirror), every rank dequantises its copy (OP_CAST_<w>_TO_F32)
//
// Message order, segment schedule and piece pipelining are those of bnet_tring_allreduce above.
namespace {

struct LocalJob {
  uint64_t value = 0;
  int nchunks = 0;
  int slot = -1;
};
constexpr int kJobSlots = 128;

int job_submit(BnetTRing* r, int cuda_dev, int slot, uint32_t op, float scale, const void* src, void* dst, size_t src_bytes,
               LocalJob* j) {
  j->slot = slot;
  j->value = ++r->jseq;     // never repeats: a stale word of the slot's previous job cannot be mistaken for completion
  volatile uint64_t* fh = r->jflags + (size_t)j->slot * cuda::kMaxChunksPerJob;
  uint64_t* fd = r->jflags_dev + (size_t)j->slot * cuda::kMaxChunksPerJob;
  return cuda::exec_transfer(cuda_dev, op, scale, src, dst, src_bytes, fh, fd, j->value, &j->nchunks);
}

bool job_done(const BnetTRing* r, const LocalJob& j) {
  if (j.slot < 0) return true;
  const volatile uint64_t* fh = r->jflags + (size_t)j.slot * cuda::kMaxChunksPerJob;
  for (int c = 0; c < j.nchunks; c++)
    if (fh[c] != j.value) return false;
  return true;
}

}  // namespace

// The wire mirror: count * (2 for bf16, 1 for fp8) bytes, device memory on the NVLink transport (the previous rank's kernels
// write it), any registered memory otherwise.  `ptr_type` is NCCL_PTR_HOST (1) or NCCL_PTR_CUDA (2).
BNET_API int bnet_tring_register_wire(BnetTRing* r, void* wire, size_t bytes, int ptr_type) {
  if (!r || !r->send_next || !r->recv_prev || !wire) return -1;
  if (r->mh_wsend) { r->send_next->dereg_mr(r->mh_wsend); r->mh_wsend = nullptr; }
  if (r->mh_wrecv) { r->recv_prev->dereg_mr(r->mh_wrecv); r->mh_wrecv = nullptr; }
  int st = r->send_next->reg_mr(wire, bytes, ptr_type, &r->mh_wsend);
  if (st) return fail(r, "regMr of the wire mirror (send side) failed: %s", status_str(st));
  st = r->recv_prev->reg_mr(wire, bytes, ptr_type, &r->mh_wrecv);
  if (st) return fail(r, "regMr of the wire mirror (receive side) failed: %s", status_str(st));
  r->wire_base = (char*)wire;
  r->wire_bytes = bytes;
  if (!r->jflags) {
    void* dp = nullptr;
    r->jflags = (volatile uint64_t*)cuda::host_alloc_mapped(sizeof(uint64_t) * kJobSlots * cuda::kMaxChunksPerJob, &dp);
    r->jflags_dev = (uint64_t*)dp;
    if (!r->jflags) return fail(r, "no memory for the executor completion words");
  }
  return 0;
}

// In-place sum all-reduce of `count` fp32 elements at `buf` with `wire_fmt` on the wire: 1 = bf16, 2 = fp8 e4m3, 3 = fp8 e5m2.
// `scale` multiplies before quantisation (fp8 only; pick it so that the PARTIAL SUMS stay inside the format's range: 448 for
// e4m3, 57344 for e5m2), its inverse is applied on the way back.  fused != 0 uses the transport's fused isend for the
// reduce-scatter hops (NVLink transport; `buf` must be registered with bnet_tring_register); fused == 0 works on any transport.
BNET_API int bnet_tring_allreduce_compressed(BnetTRing* r, void* buf, size_t count, int wire_fmt, float scale, int fused,
                                             size_t piece_elems, int inflight, int timeout_ms) {
  if (!r || !r->mh_wsend || !r->mh_wrecv || !r->jflags) return -1;
  if (wire_fmt < 1 || wire_fmt > 3) return fail(r, "wire format %d (1 = bf16, 2 = e4m3, 3 = e5m2)", wire_fmt);
  if (!(scale > 0.f)) return fail(r, "scale must be positive");
  const size_t wes = wire_fmt == 1 ? 2 : 1;
  if (wire_fmt == 1) scale = 1.0f;                                      // the bf16 passes do not scale
  const uint32_t cast_op = wire_fmt == 1 ? cuda::OP_CAST_F32_TO_BF16 : wire_fmt == 2 ? cuda::OP_CAST_F32_TO_E4M3 : cuda::OP_CAST_F32_TO_E5M2;
  const uint32_t acc_op = wire_fmt == 1 ? cuda::OP_ACC_BF16_TO_F32 : wire_fmt == 2 ? cuda::OP_ACC_E4M3_TO_F32 : cuda::OP_ACC_E5M2_TO_F32;
  const uint32_t dec_op = wire_fmt == 1 ? cuda::OP_CAST_BF16_TO_F32 : wire_fmt == 2 ? cuda::OP_CAST_E4M3_TO_F32 : cuda::OP_CAST_E5M2_TO_F32;
  const float inv = 1.0f / scale;
  char* base = (char*)buf;
  char* wire = r->wire_base;
  if (count * wes > r->wire_bytes) return fail(r, "the wire mirror holds %zu bytes, %zu needed", r->wire_bytes, count * wes);
  if ((((uintptr_t)base) & 255) != 0 || (((uintptr_t)wire) & 63) != 0) return fail(r, "buffers must be 256-byte (data) / 64-byte (wire) aligned");
  if (fused && (!r->mh_send || base < r->reg_base || base + count * 4 > r->reg_base + r->reg_bytes))
    return fail(r, "fused mode needs the data buffer registered (bnet_tring_register)");
  int cuda_dev = 0;
  if (!cuda::fake()) {
    if (!cuda::available() || !cuda::pointer_is_device(buf, &cuda_dev)) return fail(r, "the data buffer must be device memory");
  }
  const int n = r->world;
  if (piece_elems < 4096) piece_elems = 4096;
  piece_elems = piece_elems / 64 * 64;
  if (inflight < 1) inflight = 1;
  if (inflight > kMaxRequests / 2) inflight = kMaxRequests / 2;
  size_t seg = (count + n - 1) / n;
  seg = (seg + 63) / 64 * 64;                                           // elements; keeps every cut 64-byte aligned in every format
  const size_t np = (seg + piece_elems - 1) / piece_elems;
  auto piece = [&](int segment, size_t p, size_t* e0, size_t* len) {
    size_t s0 = (size_t)segment * seg, s1 = s0 + seg;
    if (s0 > count) s0 = count;
    if (s1 > count) s1 = count;
    size_t a = s0 + p * piece_elems, b = a + piece_elems;
    if (a > s1) a = s1;
    if (b > s1) b = s1;
    *e0 = a;
    *len = b - a;
  };
  const int steps = 2 * (n - 1);
  const size_t M = (size_t)steps * np;
  auto send_seg = [&](int g) { return g < n - 1 ? ((r->rank - g) % n + n) % n : ((r->rank + 1 - (g - (n - 1))) % n + n) % n; };
  auto recv_seg = [&](int g) { return g < n - 1 ? ((r->rank - g - 1) % n + n) % n : ((r->rank - (g - (n - 1))) % n + n) % n; };
  // receive side: 0 not posted, 1 posted, 2 arrived (local pass not submitted yet), 3 local pass running, 4 ready
  // send side   : 0 waiting for its input, 1 local quantise pass running, 2 quantised (or nothing to do), 3 posted, 4 done
  std::vector<Request*> rreq(M, nullptr), sreq(M, nullptr);
  std::vector<unsigned char> rst(M, 0), sst(M, 0);
  std::vector<LocalJob> rjob(M), sjob(M), ojob(np);   // ojob: the owner's own dequantise pass of all-gather step 0
  std::vector<unsigned char> ost(np, 0);              // 0 not started, 1 running, 2 done
  size_t posted_r = 0, posted_s = 0, done_r = 0, done_s = 0, ready_r = 0, owner_done = 0;
  std::vector<int> free_slots;                        // completion-word slots of the executor passes (jobs finish out of order)
  for (int i = kJobSlots - 1; i >= 0; i--) free_slots.push_back(i);
  r->last_msgs = 0;
  r->last_bytes_sent = 0;
  const uint64_t t0 = now_ns();
  CollSpan span(r->rank, ++r->op_seq, (uint64_t)count * 4);
  IdleBackoff backoff;
  auto submit = [&](uint32_t op, float sc, const void* src, void* dst, size_t src_bytes, LocalJob* j) -> int {
    if (src_bytes == 0) { j->slot = -1; return 0; }
    if (free_slots.empty()) return 1;                                   // try again later
    const int slot = free_slots.back();
    if (job_submit(r, cuda_dev, slot, op, sc, src, dst, src_bytes, j) != 0) return -1;
    free_slots.pop_back();
    return 0;
  };
  auto retire = [&](LocalJob& j) {
    if (j.slot >= 0) free_slots.push_back(j.slot);
    j.slot = -1;
  };
  const bool send_after_own_send = orders_receives_after_own_sends(r);
  while (done_s < M || ready_r < M || owner_done < np) {
    bool moved = false;
    // ---- receives: posted in message order into the wire mirror
    while (posted_r < M && posted_r - done_r < (size_t)inflight) {
      const int g = (int)(posted_r / np);
      if (send_after_own_send && g >= n - 1 && sst[posted_r - (size_t)(n - 1) * np] != 4) break;   // (as in the plain ring)
      size_t e0, len;
      piece(recv_seg(g), posted_r % np, &e0, &len);
      Request* q = nullptr;
      int st = r->recv_prev->irecv(wire + e0 * wes, len * wes, 0, r->mh_wrecv, &q);
      if (st) return fail(r, "irecv failed: %s", status_str(st));
      if (!q) break;
      rreq[posted_r] = q;
      rst[posted_r++] = 1;
      moved = true;
    }
    // ---- arrivals, then the local pass of each arrived piece (accumulate in the reduce-scatter, dequantise in the all-gather)
    for (size_t m = done_r; m < posted_r; m++) {
      if (rst[m] != 1) continue;
      int done = 0;
      size_t sz = 0;
      int st = r->recv_prev->test(rreq[m], &done, &sz);
      if (st) return fail(r, "receive %zu failed: %s", m, status_str(st));
      if (done) { rst[m] = 2; moved = true; }
    }
    while (done_r < posted_r && rst[done_r] >= 2) done_r++;
    for (size_t m = ready_r; m < done_r; m++) {
      if (rst[m] == 2) {
        const int g = (int)(m / np);
        size_t e0, len;
        piece(recv_seg(g), m % np, &e0, &len);
        int st = submit(g < n - 1 ? acc_op : dec_op, inv, wire + e0 * wes, base + e0 * 4, len * wes, &rjob[m]);
        if (st < 0) return fail(r, "executor pass failed");
        if (st == 0) { rst[m] = 3; moved = true; }
      }
      if (rst[m] == 3 && job_done(r, rjob[m])) {
        retire(rjob[m]);
        rst[m] = 4;
        moved = true;
      }
    }
    while (ready_r < M && rst[ready_r] == 4) ready_r++;
    // ---- sends: inputs ready -> (quantise) -> post in message order
    for (size_t m = posted_s; m < M && m < posted_s + (size_t)inflight; m++) {
      const int g = (int)(m / np);
      const size_t p = m % np;
      if (sst[m] == 0) {
        // input of message (g, p): reduce-scatter / first all-gather step: the accumulated piece of the previous step;
        // later all-gather steps: the narrow piece that ARRIVED in the previous step (forwarded as it is)
        if (g > 0) {
          const unsigned char need = g <= n - 1 ? 4 : 2;
          if (rst[(size_t)(g - 1) * np + p] < need) break;             // later messages depend on later arrivals: stop looking
        }
        size_t e0, len;
        piece(send_seg(g), p, &e0, &len);
        const bool quantise = (g < n - 1 && !fused) || g == n - 1;      // into the own mirror
        if (quantise) {
          int st = submit(cast_op, scale, base + e0 * 4, wire + e0 * wes, len * 4, &sjob[m]);
          if (st < 0) return fail(r, "executor pass failed");
          if (st > 0) break;
          sst[m] = 1;
        } else {
          sst[m] = 2;
        }
        moved = true;
      }
      if (sst[m] == 1 && job_done(r, sjob[m])) {
        retire(sjob[m]);
        sst[m] = 2;
        moved = true;
      }
    }
    while (posted_s < M && sst[posted_s] == 2 && posted_s - done_s < (size_t)inflight) {
      const int g = (int)(posted_s / np);
      size_t e0, len;
      piece(send_seg(g), posted_s % np, &e0, &len);
      Request* q = nullptr;
      int st;
      if (g < n - 1 && fused && len > 0) {
        st = r->send_next->isend_op(base + e0 * 4, len * 4, 0, r->mh_send, cast_op, scale, &q);
      } else {
        st = r->send_next->isend(wire + e0 * wes, len * wes, 0, r->mh_wsend, &q);
      }
      if (st) return fail(r, "isend failed: %s", status_str(st));
      if (!q) break;
      sreq[posted_s] = q;
      sst[posted_s++] = 3;
      r->last_msgs++;
      r->last_bytes_sent += len * wes;
      moved = true;
    }
    for (size_t m = done_s; m < posted_s; m++) {
      if (sst[m] != 3) continue;
      int done = 0;
      size_t sz = 0;
      int st = r->send_next->test(sreq[m], &done, &sz);
      if (st) return fail(r, "send %zu failed: %s", m, status_str(st));
      if (done) { sst[m] = 4; moved = true; }
    }
    while (done_s < posted_s && sst[done_s] == 4) done_s++;
    // ---- the owner's own copy of its finished segment goes through the same quantise -> dequantise round trip
    for (size_t p = 0; p < np; p++) {
      const size_t m0 = (size_t)(n - 1) * np + p;                       // all-gather step 0, piece p
      if (ost[p] == 0 && sst[m0] >= 2) {
        size_t e0, len;
        piece(send_seg(n - 1), p, &e0, &len);
        int st = submit(dec_op, inv, wire + e0 * wes, base + e0 * 4, len * wes, &ojob[p]);
        if (st < 0) return fail(r, "executor pass failed");
        if (st == 0) { ost[p] = 1; moved = true; }
      }
      if (ost[p] == 1 && job_done(r, ojob[p])) {
        retire(ojob[p]);
        ost[p] = 2;
        owner_done++;
        moved = true;
      }
    }
    if (!cuda::fake()) cuda::exec_kick(cuda_dev);                       // launch what the executor has collected
    if (!moved && timeout_ms > 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull)
      return fail(r, "compressed all-reduce timed out: %zu/%zu receives ready, %zu/%zu sends done", ready_r, M, done_s, M);
    backoff.step(moved);
  }
  span.ok = true;
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
  if (r->mh_wsend && r->send_next) r->send_next->dereg_mr(r->mh_wsend);
  if (r->mh_wrecv && r->recv_prev) r->recv_prev->dereg_mr(r->mh_wrecv);
  // (the job flags are deliberately not freed: cudaFreeHost waits for running kernels)
  delete r->send_next;
  delete r->recv_prev;
  delete r->listen;
  delete r;
}
