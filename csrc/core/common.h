// bnet core utilities: logging, environment knobs, small OS helpers.
//
// Behavioural parity with the reference's L6 utilities (reference:
// src/utils.rs:7-23 link speed, :132-178 IO loops, :180-198 URL parse,
// :200-205 chunk sizing) — re-designed: no busy yield loops, no panics.
#pragma once

#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <netinet/in.h>
#include <sys/socket.h>

#include "bnet/nccl_net_abi.h"

namespace bnet {

// ---- status codes used inside the engine; mapped to ncclResult_t by the shims
enum Status : int {
  kOk = 0,
  kErrSystem = 1,    // syscall / socket failure            -> ncclSystemError
  kErrInternal = 2,  // logic error / invalid state         -> ncclInternalError
  kErrRemote = 3,    // peer closed / peer reported failure -> ncclRemoteError
  kErrInvalid = 4,   // bad argument from the caller        -> ncclInvalidArgument
  kErrCuda = 5,      // CUDA failure                        -> ncclUnhandledCudaError
  kErrTimeout = 6,   // watchdog fired                      -> ncclSystemError
};
ncclResult_t to_nccl(int st);
const char* status_str(int st);

// ---- logging ---------------------------------------------------------------
// Messages go to NCCL's logger when the plugin was initialised by NCCL
// (reference: cc/v4/nccl_net_v4.cc:13-16) and to stderr when BNET_LOG_LEVEL asks.
enum LogLevel { LOG_NONE = 0, LOG_WARN = 1, LOG_INFO = 2, LOG_DEBUG = 3, LOG_TRACE = 4 };
void log_set_nccl_logger(ncclDebugLogger_t fn);
int log_level();
void log_msg(int level, const char* file, int line, const char* fmt, ...)
    __attribute__((format(printf, 4, 5)));
#define BNET_WARN(...) ::bnet::log_msg(::bnet::LOG_WARN, __FILE__, __LINE__, __VA_ARGS__)
#define BNET_INFO(...) ::bnet::log_msg(::bnet::LOG_INFO, __FILE__, __LINE__, __VA_ARGS__)
#define BNET_DEBUG(...)                                                         \
  do {                                                                          \
    if (::bnet::log_level() >= ::bnet::LOG_DEBUG)                               \
      ::bnet::log_msg(::bnet::LOG_DEBUG, __FILE__, __LINE__, __VA_ARGS__);      \
  } while (0)
#define BNET_TRACE(...)                                                         \
  do {                                                                          \
    if (::bnet::log_level() >= ::bnet::LOG_TRACE)                               \
      ::bnet::log_msg(::bnet::LOG_TRACE, __FILE__, __LINE__, __VA_ARGS__);      \
  } while (0)

// ---- environment -------------------------------------------------------------
// Every knob is looked up as BNET_<name> first and BAGUA_NET_<name> second so the
// reference's variable names keep working (reference env table: SURVEY.md §2.7).
const char* env_raw(const char* suffix);                 // BNET_x / BAGUA_NET_x
std::string env_str(const char* suffix, const char* dflt);
long long env_int(const char* suffix, long long dflt);   // malformed -> dflt (+warn)
const char* env_plain(const char* name);                 // exact name (NCCL_*, RANK)
long long env_plain_int(const char* name, long long dflt);

struct Config {
  std::string implement;     // BASIC | TOKIO(=ASYNC)
  int nstreams;              // data streams / device clusters per connection
  size_t min_chunksize;      // lower bound for a chunk
  int async_workers;         // BAGUA_NET_TOKIO_WORKER_THREADS analogue
  int rank;                  // RANK, -1 when unset
  int nvl;                   // intra-host shared-memory / NVLink transport on?
  int gdr;                   // advertise NCCL_PTR_CUDA
  int wire_compat;           // speak the reference's bare wire format
  int timeout_ms;            // watchdog for stuck transfers (0 = off)
  int spin_us;               // spin budget before a worker sleeps
  size_t shm_ring_bytes;     // per-connection bounce ring for host<->host NVL path
  std::string fault;         // BNET_FAULT_INJECT spec
  static const Config& get();   // parsed once
  static void reload();         // tests only
};

// ---- time / hashing ----------------------------------------------------------
uint64_t now_ns();
uint64_t fnv1a(const void* p, size_t n, uint64_t seed = 1469598103934665603ull);
uint64_t host_hash();         // hostname + boot id; equal <=> same OS instance
uint64_t random_u64();

// ---- call breadcrumbs -----------------------------------------------------------
// Where is every thread that is inside the plugin right now?  A CallScope costs two relaxed stores; the
// watchdog prints the scopes that have been open for too long (a proxy thread stuck in a CUDA launch behind a
// device-synchronising call of the application shows up here, not in any request state).
struct CallScope {
  explicit CallScope(const char* what);
  ~CallScope();
  int slot_, depth_;
};
// prints every open scope older than `older_than_ns` to stderr; returns how many it printed
int calltrace_dump(uint64_t older_than_ns);

// ---- chunking (reference: src/utils.rs:200-205) ---------------------------------
inline size_t chunk_size(size_t total, size_t min_chunksize, size_t expected_nchunks) {
  if (expected_nchunks == 0) expected_nchunks = 1;
  size_t c = (total + expected_nchunks - 1) / expected_nchunks;
  return c < min_chunksize ? min_chunksize : c;
}
inline size_t chunk_count(size_t total, size_t csize) { return csize ? (total + csize - 1) / csize : 0; }

// ---- "[user:pass@]host:port" (reference: src/utils.rs:180-198) --------------------
struct UserPassAddr {
  std::string user, pass, addr;
};
bool parse_user_pass_and_addr(const std::string& raw, UserPassAddr* out);
std::string base64(const std::string& in);

// ---- socket helpers -------------------------------------------------------------
union SockAddr {
  sockaddr sa;
  sockaddr_in in4;
  sockaddr_in6 in6;
};
socklen_t sockaddr_len(const SockAddr& a);
std::string sockaddr_str(const SockAddr& a);            // "127.0.0.1:8123" / "[::1]:80"
bool sockaddr_parse(const std::string& s, SockAddr* out);  // inverse of the above
int set_nonblocking(int fd, bool on);
int set_nodelay(int fd);

// Non-blocking IO that sleeps in poll() instead of spinning with yield
// (reference hot loops: src/utils.rs:132-178).  `abort` lets close() wake a
// blocked worker; timeout_ms<=0 waits forever.  Returns kOk, kErrRemote on
// EOF/reset, kErrSystem otherwise, kErrTimeout when the watchdog fires.
int write_all(int fd, const void* buf, size_t n, const std::atomic<bool>* abort, int timeout_ms);
int read_exact(int fd, void* buf, size_t n, const std::atomic<bool>* abort, int timeout_ms);

inline uint64_t be64(uint64_t v) { return __builtin_bswap64(v); }
inline uint32_t be32(uint32_t v) { return __builtin_bswap32(v); }

// sysfs
int net_if_speed_mbps(const std::string& ifname);   // default 10000 (reference: utils.rs:7-23)
std::string net_if_pci_path(const std::string& ifname);

}  // namespace bnet
