// Transport contract between the plugin shims / engine and the data movers.
//
// This is the C++ counterpart of the reference's `trait Net`
// (reference: src/interface.rs:34-74): listen / connect / accept /
// isend / irecv / test / close, plus what the reference leaves out —
// regMr with CUDA pointers, iflush, tags, error codes on every request.
//
// Design differences on purpose:
//   * comms are objects with their own fixed request pool: no global mutex
//     (reference src/lib.rs:15), no heap allocation or leak per message
//     (reference cc/bagua_net.cc:80,88)
//   * connect/accept are incremental so NCCL >= 2.1x proxies never deadlock
#pragma once

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <string>

#include "core/common.h"

namespace bnet {

constexpr int kMaxRequests = 64;      // per comm; >= NCCL_NET_MAX_REQUESTS (32)
constexpr int kMaxGroupRecvs = 8;     // receives one grouped irecv may carry (NCCL_NET_MAX_RECVS / NCCL_PROXY_MAX_SUBS)
constexpr uint32_t kHandleMagic = 0x424E4554u;  // "BNET"
constexpr uint16_t kWireVersion = 1;

// What listen() writes into NCCL's opaque handle buffer.  The first bytes are a
// plain sockaddr like the reference's handle (src/lib.rs:121-124,157-158) but a
// full sockaddr_in6 fits (the reference truncates it to 16 bytes).  64 bytes in
// total so that it also fits the v3/v4 handle size.
struct alignas(8) Handle {
  SockAddr addr;        // 28 bytes used
  uint32_t magic;       // kHandleMagic, 0 in wire-compat mode
  uint16_t version;
  uint16_t flags;       // HandleFlags
  uint64_t host_hash;   // same value <=> same OS instance
  uint64_t listen_nonce;
  uint32_t pid;
  int32_t cuda_dev;
};
static_assert(sizeof(Handle) <= 64, "handle must fit NCCL_NET_HANDLE_MAXSIZE of ncclNet v4");
enum HandleFlags : uint16_t { HF_NVL = 1, HF_ASYNC = 2 };

// First bytes on every TCP connection of a comm (reference: 8-byte BE stream id,
// nthread_…:327,376; we append the parameters both ends must agree on so a
// mismatch of NSTREAMS / MIN_CHUNKSIZE can no longer corrupt data silently).
struct Preamble {
  uint64_t stream_id_be;   // == nstreams for the control stream
  uint32_t magic;
  uint16_t version;
  uint16_t nstreams;
  uint64_t conn_nonce;     // groups the sockets of one connect()
  uint32_t impl;           // 0 = BASIC, 1 = ASYNC
  uint32_t min_chunksize;
};
static_assert(sizeof(Preamble) == 32, "preamble layout");

struct Comm;

struct MemHandle {
  void* addr = nullptr;
  size_t size = 0;
  int type = NCCL_PTR_HOST;
  uint32_t id = 0;
  int dev = -1;             // CUDA device of the buffer (staged copies run on worker threads)
  Comm* owner = nullptr;
  void* priv = nullptr;     // transport specific (NVL: export record)
  int cma = 0;              // NVL sender: 1 = other processes can read this buffer with process_vm_readv
};

enum ReqKind : uint8_t { REQ_SEND = 0, REQ_RECV = 1, REQ_FLUSH = 2 };

// Completion accounting follows the reference's RequestState
// (nthread_…:541-546,604-616): a request is done when every sub-task that was
// announced has completed; the poster's own sub-task is announced up front so
// `done` cannot be observed early.
struct Request {
  std::atomic<uint32_t> in_use{0};
  Comm* comm = nullptr;
  ReqKind kind = REQ_SEND;
  std::atomic<uint32_t> nsub{1}, ndone{0};
  std::atomic<uint64_t> nbytes{0};
  std::atomic<int> err{0};
  void* buf = nullptr;
  size_t size = 0;
  int tag = 0;
  MemHandle* mh = nullptr;
  uint64_t id = 0;
  uint64_t span = 0;
  uint64_t t_post = 0;
  uint64_t nvtx = 0;
  bool nvtx_open = false;
  void* prof_event = nullptr;   // event opened in NCCL's profiler (ncclNet v10), closed when test() reports done
  uint8_t prof_path = 0;        // BNET_PROF_PATH_* for the stop record
  uint64_t u[6] = {0, 0, 0, 0, 0, 0};   // transport scratch

  // ndone FIRST, nsub second (two sequenced statements): a dispatcher may still be announcing sub-tasks while
  // workers complete them; reading nsub first could pair a stale nsub with a later ndone and report "done" early
  bool complete() const {
    const uint32_t d = ndone.load(std::memory_order_acquire);
    const uint32_t n = nsub.load(std::memory_order_acquire);
    return d == n;
  }
  void fail(int st) {
    int z = 0;
    err.compare_exchange_strong(z, st);
  }
};

struct Comm {
  enum Kind { LISTEN, SEND, RECV };
  Kind kind;
  int dev = 0;
  uint64_t id = 0;
  std::atomic<int> broken{0};         // sticky error: later requests fail fast
  std::atomic<uint64_t> next_req{1};
  std::atomic<uint32_t> next_mr{1};
  Request pool[kMaxRequests];

  explicit Comm(Kind k);
  virtual ~Comm();
  virtual const char* transport() const = 0;

  virtual int reg_mr(void* data, size_t size, int type, MemHandle** out);
  virtual int dereg_mr(MemHandle* mh);
  virtual int isend(const void* data, size_t size, int tag, MemHandle* mh, Request** out);
  // Extension (the north star's "fused isend"): like isend, but the receiver's posted buffer is combined with the
  // payload instead of overwritten — `op` is an ExecOp (csrc/cuda/exec_body.cuh): accumulate (K4), cast (K5), both.
  // Only the device path can do that (registered CUDA buffers on both sides); everything else returns kErrInvalid.
  virtual int isend_op(const void* data, size_t size, int tag, MemHandle* mh, uint32_t op, float scale, Request** out);
  virtual int irecv(void* data, size_t size, int tag, MemHandle* mh, Request** out);
  virtual int iflush(void* data, size_t size, MemHandle* mh, Request** out);
  // done/size semantics of ncclNet test(); frees the request when *done
  virtual int test(Request* r, int* done, size_t* size);
  virtual void progress() {}

  Request* alloc_req(ReqKind kind, void* buf, size_t size, int tag, MemHandle* mh);
  void free_req(Request* r);
  int free_requests() const;          // slots of the pool that are free right now (a lower bound for the caller that owns the comm)
};

struct DeviceProps {
  std::string name, pci_path;
  uint64_t guid = 0;
  int ptr_support = NCCL_PTR_HOST;
  int speed_mbps = 10000;
  int port = 0;
  float latency_us = 0;
  int max_comms = 65536;
  int max_recvs = 1;
};

}  // namespace bnet
