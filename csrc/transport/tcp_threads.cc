// TCP transport "BASIC": one control stream + N data streams per connection,
// a dedicated worker thread per data stream and a per-connection dispatcher that
// frames messages and round-robins their chunks over the streams.
//
// Behavioural parity with the reference's default backend
// (reference: src/implement/nthread_per_socket_backend.rs:305-423 connect/send
// side, :425-522 accept/recv side, :524-631 isend/irecv/test; wire format in
// SURVEY.md Appendix B):
//   * ctrl <- u64_be(len) per message, then chunks of
//     max(ceil(len/N), MIN_CHUNKSIZE) bytes dealt round-robin to the data
//     streams; the round-robin cursor persists across messages so consecutive
//     small messages rotate over the streams (fairness, README.md:6)
//   * a request completes when the dispatcher finished announcing AND every
//     chunk worker reported
// Re-designed: workers sleep in poll() instead of spinning with yield; IO errors
// on data streams fail the request instead of panicking the worker (the
// reference hangs there, nthread_…:341,457); threads are joined on close;
// CUDA buffers are staged through a pinned bounce buffer when NCCL hands us
// device pointers on a connection that could not use the NVLink transport.
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <memory>
#include <thread>
#include <vector>

#include "core/engine.h"
#include "core/telemetry.h"
#include "cuda/cuda_iface.h"
#include "transport/workq.h"

namespace bnet {
namespace {

struct Chunk {
  char* p;
  size_t n;
  Request* r;
  bool cuda;
  int dev;
};

struct Fault {   // BNET_FAULT_INJECT="send_drop_after=<chunks>" | "recv_drop_after=<chunks>"
  long send_drop_after = -1, recv_drop_after = -1;
  static Fault parse() {
    Fault f;
    const std::string& s = Config::get().fault;
    size_t p;
    if ((p = s.find("send_drop_after=")) != std::string::npos) f.send_drop_after = atol(s.c_str() + p + 16);
    if ((p = s.find("recv_drop_after=")) != std::string::npos) f.recv_drop_after = atol(s.c_str() + p + 16);
    return f;
  }
};

constexpr size_t kBounceBytes = 1 << 20;

class TcpThreadsComm : public Comm {
 public:
  TcpThreadsComm(Kind k, int dev_, int ctrl, std::vector<int> fds, const ConnParams& p)
      : Comm(k), ctrl_(ctrl), fds_(std::move(fds)), p_(p), fault_(Fault::parse()) {
    dev = dev_;
    const Config& cfg = Config::get();
    spin_us_ = cfg.spin_us;
    timeout_ms_ = cfg.timeout_ms;
    set_nonblocking(ctrl_, true);
    for (int fd : fds_) {
      set_nonblocking(fd, true);
      set_nodelay(fd);
    }
    streams_.resize(fds_.size());
    for (size_t i = 0; i < fds_.size(); i++) {
      streams_[i].reset(new WorkQ<Chunk>());
      workers_.emplace_back([this, i] { kind == SEND ? send_worker(i) : recv_worker(i); });
    }
    dispatcher_ = std::thread([this] { kind == SEND ? send_dispatch() : recv_dispatch(); });
  }

  ~TcpThreadsComm() override {
    abort_.store(true);
    // wake anything blocked in poll(): FIN after queued data, so nothing already written is lost
    shutdown(ctrl_, SHUT_RDWR);
    for (int fd : fds_) shutdown(fd, SHUT_RDWR);
    inbox_.stop();
    for (auto& s : streams_) s->stop();
    if (dispatcher_.joinable()) dispatcher_.join();
    for (auto& w : workers_)
      if (w.joinable()) w.join();
    close(ctrl_);
    for (int fd : fds_) close(fd);
    for (void* b : bounce_) cuda::host_free_mapped(b);
  }

  const char* transport() const override { return "tcp-threads"; }

  int reg_mr(void* data, size_t size, int type, MemHandle** out) override {
    if (type == NCCL_PTR_CUDA && !Engine::get().cuda_ok()) return kErrInvalid;
    if (type != NCCL_PTR_HOST && type != NCCL_PTR_CUDA) return kErrInvalid;
    MemHandle* mh = new MemHandle;
    mh->addr = data;
    mh->size = size;
    mh->type = type;
    mh->id = next_mr.fetch_add(1);
    mh->owner = this;
    if (type == NCCL_PTR_CUDA) cuda::pointer_is_device(data, &mh->dev);
    *out = mh;
    return kOk;
  }

  int isend(const void* data, size_t size, int tag, MemHandle* mh, Request** out) override {
    return post(REQ_SEND, const_cast<void*>(data), size, tag, mh, out);
  }
  int irecv(void* data, size_t size, int tag, MemHandle* mh, Request** out) override {
    return post(REQ_RECV, data, size, tag, mh, out);
  }

 private:
  int post(ReqKind k, void* data, size_t size, int tag, MemHandle* mh, Request** out) {
    *out = nullptr;
    if ((k == REQ_SEND) != (kind == SEND)) return kErrInvalid;
    int b = broken.load(std::memory_order_acquire);
    if (b) return b;
    Request* r = alloc_req(k, data, size, tag, mh);
    if (!r) return kOk;  // would block: NCCL retries
    inbox_.push(r);
    *out = r;
    return kOk;
  }

  bool is_cuda(const Request* r) const { return r->mh && r->mh->type == NCCL_PTR_CUDA; }

  void fail_all(Request* first, int st) {
    broken.store(st, std::memory_order_release);
    if (first) {
      first->fail(st);
      first->ndone.fetch_add(1, std::memory_order_release);
    }
    Request* r;
    while (inbox_.try_pop(&r)) {
      r->fail(st);
      r->ndone.fetch_add(1, std::memory_order_release);
    }
  }

  // ---- send side ---------------------------------------------------------------
  void send_dispatch() {
    Request* r;
    size_t rr = 0;  // persists across messages (reference: nthread_…:393)
    const size_t ns = fds_.size();
    while (inbox_.pop(&r, spin_us_)) {
      int b = broken.load(std::memory_order_acquire);
      if (b) { fail_all(r, b); continue; }
      uint64_t hdr = be64((uint64_t)r->size);
      int st = write_all(ctrl_, &hdr, sizeof(hdr), &abort_, timeout_ms_);
      if (st != kOk) { fail_all(r, st); continue; }
      if (r->size) {
        size_t cs = chunk_size(r->size, p_.min_chunksize, ns);
        char* base = (char*)r->buf;
        bool cu = is_cuda(r);
        for (size_t off = 0; off < r->size; off += cs) {
          size_t n = r->size - off < cs ? r->size - off : cs;
          r->nsub.fetch_add(1, std::memory_order_release);
          streams_[rr]->push(Chunk{base + off, n, r, cu, cu ? r->mh->dev : -1});
          rr = (rr + 1) % ns;
        }
      }
      r->ndone.fetch_add(1, std::memory_order_release);  // the dispatcher's own sub-task
    }
    fail_all(nullptr, kErrRemote);
  }

  void send_worker(size_t i) {
    Chunk c;
    long nchunks = 0;
    char* bounce = nullptr;
    Telemetry& T = Telemetry::get();
    while (streams_[i]->pop(&c, spin_us_)) {
      int st = broken.load(std::memory_order_acquire);
      if (!st && fault_.send_drop_after >= 0 && nchunks >= fault_.send_drop_after) {
        shutdown(fds_[i], SHUT_RDWR);  // fault injection: this stream dies mid-flight
      }
      uint64_t t0 = now_ns();
      if (!st) {
        if (!c.cuda) {
          st = write_all(fds_[i], c.p, c.n, &abort_, timeout_ms_);
        } else {
          if (!bounce) bounce = get_bounce();
          if (!bounce) st = kErrCuda;
          for (size_t off = 0; off < c.n && !st; off += kBounceBytes) {
            size_t n = c.n - off < kBounceBytes ? c.n - off : kBounceBytes;
            if (cuda::memcpy_sync(bounce, c.p + off, n, c.dev) != 0) st = kErrCuda;
            else st = write_all(fds_[i], bounce, n, &abort_, timeout_ms_);
          }
        }
      }
      if (st) {
        c.r->fail(st);
        broken.store(st, std::memory_order_release);
      } else {
        c.r->nbytes.fetch_add(c.n, std::memory_order_relaxed);
        T.on_chunk_sent(c.n, now_ns() - t0);
      }
      nchunks++;
      c.r->ndone.fetch_add(1, std::memory_order_release);
    }
  }

  // ---- recv side ---------------------------------------------------------------
  void recv_dispatch() {
    Request* r;
    size_t rr = 0;
    const size_t ns = fds_.size();
    while (inbox_.pop(&r, spin_us_)) {
      int b = broken.load(std::memory_order_acquire);
      if (b) { fail_all(r, b); continue; }
      uint64_t hdr = 0;
      int st = read_exact(ctrl_, &hdr, sizeof(hdr), &abort_, timeout_ms_);
      if (st != kOk) { fail_all(r, st); continue; }
      size_t len = (size_t)be64(hdr);
      if (len > r->size) {
        BNET_WARN("irecv: peer sends %zu bytes into a %zu byte buffer", len, r->size);
        fail_all(r, kErrInvalid);
        continue;
      }
      if (len) {
        size_t cs = chunk_size(len, p_.min_chunksize, ns);  // identical split on both ends
        char* base = (char*)r->buf;
        bool cu = is_cuda(r);
        for (size_t off = 0; off < len; off += cs) {
          size_t n = len - off < cs ? len - off : cs;
          r->nsub.fetch_add(1, std::memory_order_release);
          streams_[rr]->push(Chunk{base + off, n, r, cu, cu ? r->mh->dev : -1});
          rr = (rr + 1) % ns;
        }
      }
      r->ndone.fetch_add(1, std::memory_order_release);
    }
    fail_all(nullptr, kErrRemote);
  }

  void recv_worker(size_t i) {
    Chunk c;
    long nchunks = 0;
    char* bounce = nullptr;
    Telemetry& T = Telemetry::get();
    while (streams_[i]->pop(&c, spin_us_)) {
      int st = broken.load(std::memory_order_acquire);
      if (!st && fault_.recv_drop_after >= 0 && nchunks >= fault_.recv_drop_after) shutdown(fds_[i], SHUT_RDWR);
      if (!st) {
        if (!c.cuda) {
          st = read_exact(fds_[i], c.p, c.n, &abort_, timeout_ms_);
        } else {
          if (!bounce) bounce = get_bounce();
          if (!bounce) st = kErrCuda;
          for (size_t off = 0; off < c.n && !st; off += kBounceBytes) {
            size_t n = c.n - off < kBounceBytes ? c.n - off : kBounceBytes;
            st = read_exact(fds_[i], bounce, n, &abort_, timeout_ms_);
            if (!st && cuda::memcpy_sync(c.p + off, bounce, n, c.dev) != 0) st = kErrCuda;
          }
        }
      }
      if (st) {
        c.r->fail(st);
        broken.store(st, std::memory_order_release);
      } else {
        c.r->nbytes.fetch_add(c.n, std::memory_order_relaxed);
        T.on_chunk_recv(c.n);
      }
      nchunks++;
      c.r->ndone.fetch_add(1, std::memory_order_release);
    }
  }

  char* get_bounce() {
    void* dptr = nullptr;
    void* b = cuda::host_alloc_mapped(kBounceBytes, &dptr);
    std::lock_guard<std::mutex> lk(bounce_mu_);
    if (b) bounce_.push_back(b);
    return (char*)b;
  }

  int ctrl_;
  std::vector<int> fds_;
  ConnParams p_;
  Fault fault_;
  int spin_us_ = 0, timeout_ms_ = 0;
  std::atomic<bool> abort_{false};
  WorkQ<Request*> inbox_;
  std::vector<std::unique_ptr<WorkQ<Chunk>>> streams_;
  std::vector<std::thread> workers_;
  std::thread dispatcher_;
  std::mutex bounce_mu_;
  std::vector<void*> bounce_;
};

}  // namespace

Comm* tcp_threads_make_send(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p) {
  return new TcpThreadsComm(Comm::SEND, dev, ctrl_fd, std::move(data_fds), p);
}
Comm* tcp_threads_make_recv(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p) {
  return new TcpThreadsComm(Comm::RECV, dev, ctrl_fd, std::move(data_fds), p);
}

}  // namespace bnet
