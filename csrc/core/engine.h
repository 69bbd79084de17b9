// Engine: device registry + connection establishment, shared by every ABI shim.
//
// Replaces the reference's L2 adaptor + L3 FFI + backend selection
// (reference: cc/bagua_net.cc:1-154 singleton & handle boxing,
// src/lib.rs:19-37 backend choice by BAGUA_NET_IMPLEMENT).
#pragma once

#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "core/netif.h"
#include "transport/transport.h"

namespace bnet {

struct ConnParams {
  int nstreams;
  size_t min_chunksize;
  int impl;            // 0 BASIC, 1 ASYNC
  uint64_t nonce;
  bool compat;         // bare reference wire format
};

// transport factories (each transport lives in its own translation unit)
Comm* tcp_threads_make_send(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p);
Comm* tcp_threads_make_recv(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p);
Comm* tcp_async_make_send(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p);
Comm* tcp_async_make_recv(int dev, int ctrl_fd, std::vector<int> data_fds, const ConnParams& p);
// NVL (intra-host shared memory + NVLink) — returns nullptr when it cannot be set up
Comm* nvl_connect(int dev, const Handle& h);
Comm* nvl_accept(int dev, int uds_fd, const void* hello, size_t hello_len);
size_t nvl_hello_size();
bool nvl_available();
std::string nvl_uds_name(uint32_t pid, uint64_t listen_nonce);

class ListenComm : public Comm {
 public:
  ListenComm() : Comm(LISTEN) {}
  ~ListenComm() override;
  const char* transport() const override { return "listen"; }
  int tcp_fd = -1;
  int uds_fd = -1;
  Handle handle{};
  struct Sock {   // accepted, preamble not complete yet
    int fd;
    bool uds;
    size_t got;
    unsigned char buf[256];
  };
  struct Pending {   // sockets of one connect(), grouped by nonce
    ConnParams p;
    int ctrl = -1;
    std::map<int, int> data;   // stream id -> fd (ordered like the sender's)
  };
  std::vector<Sock> socks;
  std::map<uint64_t, Pending> pending;
  std::mutex mu;
};

// ncclNet v10: events of this plugin inside NCCL's profiler (include/bnet/bnet_profiler.h)
void profiler_set_callback(ncclProfilerCallback_t fn);
void profiler_start(Request* r, void* parent_handle);   // no-op without a callback or a parent handle
void profiler_stop(Request* r);

class Engine {
 public:
  static Engine& get();
  int init();                      // idempotent
  int ndev();
  int props(int dev, DeviceProps* out);
  int listen(int dev, void* handle_out, size_t handle_cap, ListenComm** out);
  int connect(int dev, const void* handle, Comm** out);          // never waits for the peer's accept()
  int accept(ListenComm* l, Comm** out, bool blocking);          // *out == nullptr: not ready yet
  const std::vector<NetIf>& devices() const { return devs_; }
  bool cuda_ok() const { return cuda_ok_; }
  int gpu_of_dev(int dev) const { return dev >= 0 && dev < (int)gpu_of_dev_.size() ? gpu_of_dev_[dev] : -1; }

 private:
  Engine() = default;
  std::mutex mu_;
  bool inited_ = false;
  bool cuda_ok_ = false;
  std::vector<NetIf> devs_;             // per NCCL device: the interface its TCP side listens on
  std::vector<int> gpu_of_dev_;         // per NCCL device: the GPU it stands for, -1 for a plain NIC device
  std::vector<DeviceProps> props_;
};

}  // namespace bnet
