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
// Re-designed: no async runtime — edge-triggered epoll state machines, one per
// comm, pinned to one of the pool's loops; no per-message task spawning.
#include <errno.h>
#include <string.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "core/engine.h"
#include "core/telemetry.h"
#include "cuda/cuda_iface.h"

namespace bnet {
namespace {

class AsyncComm;

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
  void add(AsyncComm* c);
  void remove(AsyncComm* c);       // returns once the loop no longer touches c
  void kick(AsyncComm* c) {        // c has new requests
    {
      std::lock_guard<std::mutex> lk(mu_);
      kicked_.push_back(c);
    }
    wake();
  }

 private:
  void run();
  int ep_, ev_;
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<AsyncComm*> to_add_, to_remove_, kicked_;
  friend class AsyncComm;
};

class Pool {
 public:
  static Pool& get() {
    static Pool* p = new Pool();  // leaked: loops must outlive static destruction order
    return *p;
  }
  Loop* pick() { return loops_[next_.fetch_add(1) % loops_.size()]; }

 private:
  Pool() {
    int n = Config::get().async_workers;
    for (int i = 0; i < n; i++) loops_.push_back(new Loop());
  }
  std::vector<Loop*> loops_;
  std::atomic<size_t> next_{0};
};

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
    ch_.resize(fds_.size());
    loop_ = Pool::get().pick();
    loop_->add(this);
  }
  ~AsyncComm() override {
    loop_->remove(this);
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

  // ---- everything below runs on the loop thread ------------------------------------
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
        size_t cs = len ? chunk_size(len, p_.min_chunksize, fds_.size()) : 0;
        left_ = 0;
        for (size_t j = 0; j < ch_.size(); j++) {
          size_t off = j * cs;
          if (len && off < len) {
            ch_[j].p = io_base_ + off;
            ch_[j].n = len - off < cs ? len - off : cs;
            ch_[j].t0 = now_ns();
            left_++;
          } else {
            ch_[j].p = nullptr;
            ch_[j].n = 0;
          }
          ch_[j].total = ch_[j].n;
        }
        st_ = DATA;
      }
      if (st_ == DATA) {
        bool blocked = false;
        for (size_t j = 0; j < ch_.size() && st_ == DATA; j++) {
          Ch& c = ch_[j];
          while (c.n) {
            ssize_t n = kind == SEND ? ::send(fds_[j], c.p, c.n, MSG_NOSIGNAL) : ::recv(fds_[j], c.p, c.n, 0);
            if (n > 0) {
              c.p += n;
              c.n -= (size_t)n;
              if (!c.n) {
                left_--;
                Telemetry& T = Telemetry::get();
                if (kind == SEND) T.on_chunk_sent(c.total, now_ns() - c.t0); else T.on_chunk_recv(c.total);
              }
              continue;
            }
            if (n < 0 && errno == EINTR) continue;
            if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) { blocked = true; break; }
            finish(n == 0 || errno == ECONNRESET || errno == EPIPE ? kErrRemote : kErrSystem);
            break;
          }
        }
        if (st_ != DATA) continue;
        if (left_ == 0) {
          int st = kOk;
          if (kind == RECV && cuda_cur_ && len_ && cuda::memcpy_sync(cur_->buf, stage_.data(), len_, cur_->mh->dev) != 0) st = kErrCuda;
          cur_->nbytes.store(len_, std::memory_order_relaxed);
          finish(st);
          continue;
        }
        if (blocked) return;
      }
    }
  }

  void on_removed() {   // loop thread: fail whatever is still queued
    dead_ = true;
    if (cur_) finish(kErrRemote);
    std::lock_guard<std::mutex> lk(in_mu_);
    for (Request* r : inbox_) {
      r->fail(kErrRemote);
      r->ndone.fetch_add(1, std::memory_order_release);
    }
    inbox_.clear();
  }

  int ctrl_;
  std::vector<int> fds_;

 private:
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
    loop_->kick(this);
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

  struct Ch {
    char* p = nullptr;
    size_t n = 0, total = 0;
    uint64_t t0 = 0;
  };
  enum State { IDLE, HDR, DATA };
  ConnParams p_;
  Loop* loop_;
  std::mutex in_mu_;
  std::deque<Request*> inbox_;
  // loop-thread state
  State st_ = IDLE;
  Request* cur_ = nullptr;
  uint32_t hdr_ = 0;
  size_t hdr_off_ = 0, len_ = 0;
  std::vector<Ch> ch_;
  size_t left_ = 0;
  char* io_base_ = nullptr;
  bool cuda_cur_ = false, dead_ = false;
  std::vector<char> stage_;
};

void Loop::add(AsyncComm* c) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    to_add_.push_back(c);
  }
  wake();
}

void Loop::remove(AsyncComm* c) {
  std::unique_lock<std::mutex> lk(mu_);
  to_remove_.push_back(c);
  lk.unlock();
  wake();
  lk.lock();
  // the loop erases c from the list once it has detached it and failed its requests
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
    bool woke = false;
    for (int i = 0; i < n; i++) {
      if (evs[i].data.ptr == nullptr) { woke = true; continue; }
    }
    if (woke) {
      uint64_t v;
      while (read(ev_, &v, sizeof(v)) > 0) {}
    }
    std::vector<AsyncComm*> add, rem, kick;
    {
      std::lock_guard<std::mutex> lk(mu_);
      add.swap(to_add_);
      rem = to_remove_;
      kick.swap(kicked_);
    }
    for (AsyncComm* c : add) {
      epoll_event e{};
      e.events = EPOLLIN | EPOLLOUT | EPOLLET | EPOLLRDHUP;
      e.data.ptr = c;
      epoll_ctl(ep_, EPOLL_CTL_ADD, c->ctrl_, &e);
      for (int fd : c->fds_) epoll_ctl(ep_, EPOLL_CTL_ADD, fd, &e);
    }
    auto removed = [&](AsyncComm* c) {
      for (AsyncComm* x : rem)
        if (x == c) return true;
      return false;
    };
    for (AsyncComm* c : rem) {
      epoll_ctl(ep_, EPOLL_CTL_DEL, c->ctrl_, nullptr);
      for (int fd : c->fds_) epoll_ctl(ep_, EPOLL_CTL_DEL, fd, nullptr);
      c->on_removed();
    }
    for (int i = 0; i < n; i++) {
      AsyncComm* c = (AsyncComm*)evs[i].data.ptr;
      if (c && !removed(c)) c->on_ready();
    }
    for (AsyncComm* c : kick)
      if (!removed(c)) c->on_ready();
    if (!rem.empty()) {
      std::lock_guard<std::mutex> lk(mu_);
      for (AsyncComm* c : rem) {
        for (size_t i = 0; i < to_remove_.size(); i++)
          if (to_remove_[i] == c) { to_remove_.erase(to_remove_.begin() + i); break; }
        // a kick that raced with the removal must not reach a freed comm
        for (size_t i = 0; i < kicked_.size();)
          if (kicked_[i] == c) kicked_.erase(kicked_.begin() + i); else i++;
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
