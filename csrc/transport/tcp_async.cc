// TCP transport "TOKIO"/ASYNC: the same multi-stream idea driven by a small pool
// of epoll event loops instead of a thread per stream.
//
// Behavioural parity with the reference's async backend
// (reference: src/implement/tokio_backend.rs:336-476 connect, :478-592 accept,
// :372-416 datapass task, :447-472 ctrl task):
//   * ctrl <- u32_be(len) per message (4-byte header, not 8 like BASIC)
//   * at most N chunks per message, chunk j always travels on stream j
//   * message-serial: message k+1 starts only when message k finished on all
//     streams; zero-length messages complete right after the header
//   * worker count from BAGUA_NET_TOKIO_WORKER_THREADS, default MIN_CHUNKSIZE 65535
// Re-designed: no async runtime and no per-message task spawning — edge-triggered epoll
// state machines.  Every comm has a HOME loop that owns all of its sockets and the message
// state machine.  The chunks of a LARGE message are lent to other loops of the pool for the
// duration of the chunk, so that several threads copy at once (the reference joins its
// per-stream futures inside one task, i.e. on one thread at a time); small messages never
// leave the home loop.
#include <errno.h>
#include <string.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "core/engine.h"
#include "core/telemetry.h"
#include "cuda/cuda_iface.h"

namespace bnet {
namespace {

class AsyncComm;

// what an epoll event (or a kick) refers to: the comm's message state machine (j < 0) or one data stream
struct Ctx {
  AsyncComm* comm;
  int j;
};

class Loop {
 public:
  Loop() {
    ep_ = epoll_create1(EPOLL_CLOEXEC);
    ev_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
    epoll_event e{};
    e.events = EPOLLIN;
    e.data.ptr = nullptr;  // nullptr marks the wake-up fd
    epoll_ctl(ep_, EPOLL_CTL_ADD, ev_, &e);
    th_ = std::thread([this] { run(); });
  }
  void wake() {
    uint64_t one = 1;
    ssize_t r = write(ev_, &one, sizeof(one));
    (void)r;
  }
  void add(int fd, Ctx* ctx) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      to_add_.push_back({fd, ctx});
    }
    wake();
  }
  void remove(AsyncComm* c);       // returns once the loop no longer touches c (any of its contexts)
  void kick(Ctx* ctx) {            // ctx has something to do
    {
      std::lock_guard<std::mutex> lk(mu_);
      kicked_.push_back(ctx);
    }
    wake();
  }
  bool on_this_thread() const { return std::this_thread::get_id() == th_.get_id(); }
  // temporary registration of a data stream for the duration of one large chunk (epoll_ctl is thread safe)
  void attach(int fd, Ctx* ctx) {
    epoll_event e{};
    e.events = EPOLLIN | EPOLLOUT | EPOLLET | EPOLLRDHUP;
    e.data.ptr = ctx;
    epoll_ctl(ep_, EPOLL_CTL_ADD, fd, &e);
  }
  void detach(int fd) { epoll_ctl(ep_, EPOLL_CTL_DEL, fd, nullptr); }

 private:
  void run();
  int ep_, ev_;
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_;
  struct Reg { int fd; Ctx* ctx; };
  std::vector<Reg> to_add_, regs_;   // regs_: loop thread only
  std::vector<AsyncComm*> to_remove_;
  std::vector<Ctx*> kicked_;
};

class Pool {
 public:
  static Pool& get() {
    static Pool* p = new Pool();  // leaked: loops must outlive static destruction order
    return *p;
  }
  size_t size() const { return loops_.size(); }
  size_t pick() { return next_.fetch_add(1) % loops_.size(); }
  Loop* loop(size_t i) { return loops_[i % loops_.size()]; }

 private:
  Pool() {
    int n = Config::get().async_workers;
    if (n < 1) n = 1;
    for (int i = 0; i < n; i++) loops_.push_back(new Loop());
  }
  std::vector<Loop*> loops_;
  std::atomic<size_t> next_{0};
};

// Chunks at least this large are handed to another loop of the pool for the duration of the chunk (parallel
// copies; costs two epoll_ctl calls and two cross-thread wake-ups, ~25 us); smaller ones stay on the home loop.
constexpr size_t kParallelChunk = 512 * 1024;
// from how many large chunks of one message on are they lent to other loops of the pool (BNET_ASYNC_LEND_MIN)
inline int lend_min_chunks() {
  static const int v = [] { long long n = env_int("ASYNC_LEND_MIN", 3); return n < 1 ? 1 : (int)n; }();
  return v;
}

class AsyncComm : public Comm {
 public:
  AsyncComm(Kind k, int dev_, int ctrl, std::vector<int> fds, const ConnParams& p)
      : Comm(k), ctrl_(ctrl), fds_(std::move(fds)), p_(p) {
    dev = dev_;
    set_nonblocking(ctrl_, true);
    for (int fd : fds_) {
      set_nonblocking(fd, true);
      set_nodelay(fd);
    }
    const size_t h = Pool::get().pick();
    home_ = Pool::get().loop(h);
    home_ctx_ = {this, -1};
    ch_ = std::unique_ptr<Ch[]>(new Ch[fds_.size()]);
    sctx_.resize(fds_.size());
    sloop_.resize(fds_.size());
    for (size_t j = 0; j < fds_.size(); j++) {
      sctx_[j] = {this, (int)j};
      sloop_[j] = Pool::get().loop(h + j);      // stream 0 stays on the home loop
    }
    // every socket is permanently registered with the home loop; a stream is attached to its helper loop only
    // while a large chunk is in flight on it
    home_->add(ctrl_, &home_ctx_);
    for (size_t j = 0; j < fds_.size(); j++) home_->add(fds_[j], &sctx_[j]);
  }
  ~AsyncComm() override {
    // helper loops first (nobody is inside drive() afterwards), the home loop last (it fails what is queued)
    for (size_t j = 0; j < fds_.size(); j++)
      if (ch_[j].armed.load(std::memory_order_acquire) == kHelper) sloop_[j]->detach(fds_[j]);
    std::vector<Loop*> done;
    auto once = [&](Loop* l) {
      for (Loop* d : done)
        if (d == l) return;
      done.push_back(l);
      l->remove(this);
    };
    for (Loop* l : sloop_)
      if (l != home_) once(l);
    once(home_);
    shutdown(ctrl_, SHUT_RDWR);
    close(ctrl_);
    for (int fd : fds_) {
      shutdown(fd, SHUT_RDWR);
      close(fd);
    }
  }
  const char* transport() const override { return "tcp-async"; }

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

  // ---- event entry point (any loop thread) ------------------------------------------------------------
  void on_event(Ctx* ctx, Loop* self) {
    if (ctx->j < 0) on_ready(); else on_stream(ctx->j, self);
  }
  bool is_home(const Loop* l) const { return l == home_; }

  // ---- home loop: fail whatever is still queued ---------------------------------------------------------
  void on_removed() {
    dead_ = true;
    if (cur_) finish(kErrRemote);
    std::lock_guard<std::mutex> lk(in_mu_);
    for (Request* r : inbox_) {
      r->fail(kErrRemote);
      r->ndone.fetch_add(1, std::memory_order_release);
    }
    inbox_.clear();
  }

 private:
  struct Ch {
    char* p = nullptr;
    size_t n = 0, total = 0;
    uint64_t t0 = 0;
    std::atomic<int> armed{0};   // who continues the chunk when the socket becomes ready: kIdle / kHome / kHelper
  };
  enum { kIdle = 0, kHome = 1, kHelper = 2 };
  enum IoResult { IO_DONE, IO_BLOCKED, IO_ERROR };

  // move bytes of chunk j until it is done or the socket would block; runs on whichever thread owns stream j
  IoResult drive(int j, int* err) {
    Ch& c = ch_[j];
    while (c.n) {
      ssize_t n = kind == SEND ? ::send(fds_[j], c.p, c.n, MSG_NOSIGNAL) : ::recv(fds_[j], c.p, c.n, 0);
      if (n > 0) {
        c.p += n;
        c.n -= (size_t)n;
        continue;
      }
      if (n < 0 && errno == EINTR) continue;
      if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) return IO_BLOCKED;
      *err = (n == 0 || errno == ECONNRESET || errno == EPIPE) ? kErrRemote : kErrSystem;
      return IO_ERROR;
    }
    Telemetry& T = Telemetry::get();
    if (kind == SEND) T.on_chunk_sent(c.total, now_ns() - c.t0); else T.on_chunk_recv(c.total);
    return IO_DONE;
  }

  // chunk j finished (or failed) on some thread: the last one wakes the message state machine
  void chunk_finished(int err) {
    if (err) {
      int expected = 0;
      serr_.compare_exchange_strong(expected, err, std::memory_order_acq_rel);
    }
    if (left_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      if (home_->on_this_thread()) on_ready(); else home_->kick(&home_ctx_);
    }
  }

  // an event or a kick for data stream j, on the home loop or on the helper loop the chunk was handed to
  void on_stream(int j, Loop* self) {
    Ch& c = ch_[j];
    const int owner = c.armed.load(std::memory_order_acquire);
    if (owner == kIdle || (owner == kHome) != (self == home_)) return;   // not ours to continue
    int err = 0;
    IoResult r = drive(j, &err);
    if (r == IO_BLOCKED) return;                                // edge-triggered: the next event continues
    if (owner == kHelper) self->detach(fds_[j]);
    c.armed.store(kIdle, std::memory_order_release);
    chunk_finished(r == IO_ERROR ? err : 0);
  }

  // the socket would block on the home loop: its (permanent) registration there delivers the next edge
  void arm_home(int j) { ch_[j].armed.store(kHome, std::memory_order_release); }

  // hand a large chunk to the stream's helper loop: attach the socket there (reports current readiness at once)
  // and kick it (it tries the IO immediately, so no edge can be lost)
  void arm_helper(int j) {
    ch_[j].armed.store(kHelper, std::memory_order_release);
    sloop_[j]->attach(fds_[j], &sctx_[j]);
    sloop_[j]->kick(&sctx_[j]);
  }

  // home loop: the message state machine
  void on_ready() {
    if (dead_) return;
    for (;;) {
      if (st_ == IDLE) {
        {
          std::lock_guard<std::mutex> lk(in_mu_);
          if (inbox_.empty()) return;
          cur_ = inbox_.front();
          inbox_.pop_front();
        }
        int b = broken.load(std::memory_order_acquire);
        if (b) { finish(b); continue; }
        hdr_off_ = 0;
        cuda_cur_ = cur_->mh && cur_->mh->type == NCCL_PTR_CUDA;
        if (kind == SEND) {
          hdr_ = be32((uint32_t)cur_->size);
          io_base_ = (char*)cur_->buf;
          if (cuda_cur_ && cur_->size) {
            stage_.resize(cur_->size);
            if (cuda::memcpy_sync(stage_.data(), cur_->buf, cur_->size, cur_->mh->dev) != 0) { finish(kErrCuda); continue; }
            io_base_ = stage_.data();
          }
        }
        st_ = HDR;
      }
      if (st_ == HDR) {
        char* h = (char*)&hdr_;
        while (hdr_off_ < 4) {
          ssize_t n = kind == SEND ? ::send(ctrl_, h + hdr_off_, 4 - hdr_off_, MSG_NOSIGNAL)
                                   : ::recv(ctrl_, h + hdr_off_, 4 - hdr_off_, 0);
          if (n > 0) { hdr_off_ += (size_t)n; continue; }
          if (n < 0 && errno == EINTR) continue;
          if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) return;  // edge-triggered: wait for the next event
          finish(n == 0 || errno == ECONNRESET || errno == EPIPE ? kErrRemote : kErrSystem);
          break;
        }
        if (st_ != HDR) continue;
        size_t len = kind == SEND ? cur_->size : (size_t)be32(hdr_);
        if (kind == RECV) {
          if (len > cur_->size) {
            BNET_WARN("irecv(async): peer sends %zu bytes into a %zu byte buffer", len, cur_->size);
            finish(kErrInvalid);
            continue;
          }
          io_base_ = (char*)cur_->buf;
          if (cuda_cur_ && len) {
            stage_.resize(len);
            io_base_ = stage_.data();
          }
        }
        len_ = len;
        // <= nstreams chunks, chunk j on stream j (tokio_…:392-403)
        const size_t ns = fds_.size();
        size_t cs = len ? chunk_size(len, p_.min_chunksize, ns) : 0;
        int nchunks = 0;
        for (size_t j = 0; j < ns; j++) {
          size_t off = j * cs;
          Ch& c = ch_[j];
          if (len && off < len) {
            c.p = io_base_ + off;
            c.n = len - off < cs ? len - off : cs;
            c.t0 = now_ns();
            nchunks++;
          } else {
            c.p = nullptr;
            c.n = 0;
          }
          c.total = c.n;
        }
        st_ = DATA;
        serr_.store(0, std::memory_order_relaxed);
        // +1: our own reference, dropped below — the message cannot complete while we are still dealing chunks
        left_.store(nchunks + 1, std::memory_order_release);
        // lending pays off from three large chunks on (measured on loopback: with two, the single loop that
        // alternates between the sockets is as fast and needs no cross-thread wake-ups)
        int big = 0;
        for (size_t j = 0; j < ns; j++) big += ch_[j].total >= kParallelChunk;
        const bool lend = big >= lend_min_chunks();
        for (size_t j = 0; j < ns; j++) {
          Ch& c = ch_[j];
          if (!c.total) continue;
          if (lend && c.total >= kParallelChunk && sloop_[j] != home_) {
            arm_helper((int)j);                // large: another loop copies it, in parallel with the others
            continue;
          }
          int err = 0;
          IoResult r = drive((int)j, &err);    // small (or our own stream): right here
          if (r == IO_BLOCKED) {
            arm_home((int)j);                  // continue when the socket is ready
            continue;
          }
          if (r == IO_ERROR) note_err(err);
          left_.fetch_sub(1, std::memory_order_acq_rel);
        }
        if (left_.fetch_sub(1, std::memory_order_acq_rel) != 1) return;   // chunks still in flight: the last one wakes us
      }
      if (st_ == DATA) {
        if (left_.load(std::memory_order_acquire) != 0) return;
        int st = serr_.load(std::memory_order_acquire);
        if (!st && kind == RECV && cuda_cur_ && len_ && cuda::memcpy_sync(cur_->buf, stage_.data(), len_, cur_->mh->dev) != 0)
          st = kErrCuda;
        if (!st) cur_->nbytes.store(len_, std::memory_order_relaxed);
        finish(st);
        continue;
      }
    }
  }

  void note_err(int err) {
    int expected = 0;
    serr_.compare_exchange_strong(expected, err, std::memory_order_acq_rel);
  }

  int post(ReqKind k, void* data, size_t size, int tag, MemHandle* mh, Request** out) {
    *out = nullptr;
    if ((k == REQ_SEND) != (kind == SEND)) return kErrInvalid;
    if (size > 0xffffffffull) return kErrInvalid;  // u32 length header
    int b = broken.load(std::memory_order_acquire);
    if (b) return b;
    Request* r = alloc_req(k, data, size, tag, mh);
    if (!r) return kOk;
    {
      std::lock_guard<std::mutex> lk(in_mu_);
      inbox_.push_back(r);
    }
    home_->kick(&home_ctx_);
    *out = r;
    return kOk;
  }

  void finish(int st) {
    if (st) {
      cur_->fail(st);
      broken.store(st, std::memory_order_release);
    }
    cur_->ndone.fetch_add(1, std::memory_order_release);  // nsub stays 1 (tokio_…:406-414)
    cur_ = nullptr;
    st_ = IDLE;
  }

  enum State { IDLE, HDR, DATA };
  int ctrl_;
  std::vector<int> fds_;
  ConnParams p_;
  Loop* home_ = nullptr;
  Ctx home_ctx_{};
  std::vector<Ctx> sctx_;
  std::vector<Loop*> sloop_;
  std::mutex in_mu_;
  std::deque<Request*> inbox_;
  std::unique_ptr<Ch[]> ch_;
  std::atomic<int> left_{0};     // chunks (+1 while the home loop deals them) of the current message still in flight
  std::atomic<int> serr_{0};     // first stream error of the current message
  // home-loop state
  State st_ = IDLE;
  Request* cur_ = nullptr;
  uint32_t hdr_ = 0;
  size_t hdr_off_ = 0, len_ = 0;
  char* io_base_ = nullptr;
  bool cuda_cur_ = false, dead_ = false;
  std::vector<char> stage_;
};

void Loop::remove(AsyncComm* c) {
  std::unique_lock<std::mutex> lk(mu_);
  to_remove_.push_back(c);
  lk.unlock();
  wake();
  lk.lock();
  // the loop erases c from the list once it has detached it (and, on its home loop, failed its requests)
  cv_.wait(lk, [&] {
    for (AsyncComm* x : to_remove_)
      if (x == c) return false;
    return true;
  });
}

void Loop::run() {
  std::vector<epoll_event> evs(256);
  for (;;) {
    int n = epoll_wait(ep_, evs.data(), (int)evs.size(), 1000);
    if (n < 0 && errno != EINTR) break;
    if (n < 0) n = 0;
    bool woke = false;
    for (int i = 0; i < n; i++) {
      if (evs[i].data.ptr == nullptr) { woke = true; continue; }
    }
    if (woke) {
      uint64_t v;
      while (read(ev_, &v, sizeof(v)) > 0) {}
    }
    std::vector<Reg> add;
    std::vector<AsyncComm*> rem;
    std::vector<Ctx*> kick;
    {
      std::lock_guard<std::mutex> lk(mu_);
      add.swap(to_add_);
      rem = to_remove_;
      kick.swap(kicked_);
    }
    for (const Reg& r : add) {
      epoll_event e{};
      e.events = EPOLLIN | EPOLLOUT | EPOLLET | EPOLLRDHUP;
      e.data.ptr = r.ctx;
      epoll_ctl(ep_, EPOLL_CTL_ADD, r.fd, &e);
      regs_.push_back(r);
    }
    auto removed = [&](AsyncComm* c) {
      for (AsyncComm* x : rem)
        if (x == c) return true;
      return false;
    };
    for (AsyncComm* c : rem) {
      for (size_t i = 0; i < regs_.size();) {
        if (regs_[i].ctx->comm == c) {
          epoll_ctl(ep_, EPOLL_CTL_DEL, regs_[i].fd, nullptr);
          regs_.erase(regs_.begin() + i);
        } else {
          i++;
        }
      }
      if (c->is_home(this)) c->on_removed();
    }
    for (int i = 0; i < n; i++) {
      Ctx* ctx = (Ctx*)evs[i].data.ptr;
      if (ctx && !removed(ctx->comm)) ctx->comm->on_event(ctx, this);
    }
    for (Ctx* ctx : kick)
      if (!removed(ctx->comm)) ctx->comm->on_event(ctx, this);
    if (!rem.empty()) {
      std::lock_guard<std::mutex> lk(mu_);
      for (AsyncComm* c : rem) {
        for (size_t i = 0; i < to_remove_.size(); i++)
          if (to_remove_[i] == c) { to_remove_.erase(to_remove_.begin() + i); break; }
        // a kick that raced with the removal must not reach a freed comm
        for (size_t i = 0; i < kicked_.size();)
          if (kicked_[i]->comm == c) kicked_.erase(kicked_.begin() + i); else i++;
      }
      cv_.notify_all();
    }
  }
}

}  // namespace

Comm* tcp_async_make_send(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p) {
  return new AsyncComm(Comm::SEND, dev, ctrl_fd, std::move(data_fds), p);
}
Comm* tcp_async_make_recv(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p) {
  return new AsyncComm(Comm::RECV, dev, ctrl_fd, std::move(data_fds), p);
}

}  // namespace bnet
