#include "core/common.h"

#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <fstream>
#include <mutex>
#include <random>

namespace bnet {

// ------------------------------------------------------------------ status
ncclResult_t to_nccl(int st) {
  switch (st) {
    case kOk: return ncclSuccess;
    case kErrSystem: return ncclSystemError;
    case kErrInternal: return ncclInternalError;
    case kErrRemote: return ncclRemoteError;
    case kErrInvalid: return ncclInvalidArgument;
    case kErrCuda: return ncclUnhandledCudaError;
    case kErrTimeout: return ncclSystemError;
    default: return ncclInternalError;
  }
}

const char* status_str(int st) {
  switch (st) {
    case kOk: return "ok";
    case kErrSystem: return "system error";
    case kErrInternal: return "internal error";
    case kErrRemote: return "remote error";
    case kErrInvalid: return "invalid argument";
    case kErrCuda: return "cuda error";
    case kErrTimeout: return "timeout";
    default: return "unknown";
  }
}

// ------------------------------------------------------------------ logging
static std::atomic<ncclDebugLogger_t> g_nccl_logger{nullptr};
static int g_log_level = -1;

void log_set_nccl_logger(ncclDebugLogger_t fn) { g_nccl_logger.store(fn); }

int log_level() {
  if (g_log_level < 0) {
    const char* e = env_raw("LOG_LEVEL");
    int lv = LOG_WARN;
    if (e) {
      if (!strcasecmp(e, "NONE")) lv = LOG_NONE;
      else if (!strcasecmp(e, "WARN")) lv = LOG_WARN;
      else if (!strcasecmp(e, "INFO")) lv = LOG_INFO;
      else if (!strcasecmp(e, "DEBUG")) lv = LOG_DEBUG;
      else if (!strcasecmp(e, "TRACE")) lv = LOG_TRACE;
      else lv = atoi(e);
    }
    g_log_level = lv;
  }
  return g_log_level;
}

void log_msg(int level, const char* file, int line, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  ncclDebugLogger_t fn = g_nccl_logger.load();
  if (fn) {
    ncclDebugLogLevel nl = level <= LOG_WARN ? NCCL_LOG_WARN : (level <= LOG_INFO ? NCCL_LOG_INFO : NCCL_LOG_TRACE);
    fn(nl, NCCL_NET | NCCL_INIT, file, line, "BNet: %s", buf);
  }
  if (level <= log_level() && (!fn || level <= LOG_WARN || env_raw("LOG_LEVEL"))) {
    const char* base = strrchr(file, '/');
    fprintf(stderr, "[bnet %s %s:%d] %s\n",
            level <= LOG_WARN ? "WARN" : level == LOG_INFO ? "INFO" : level == LOG_DEBUG ? "DEBUG" : "TRACE",
            base ? base + 1 : file, line, buf);
  }
}

// ------------------------------------------------------------------ env
const char* env_raw(const char* suffix) {
  char name[128];
  snprintf(name, sizeof(name), "BNET_%s", suffix);
  const char* v = getenv(name);
  if (v && *v) return v;
  snprintf(name, sizeof(name), "BAGUA_NET_%s", suffix);
  v = getenv(name);
  if (v && *v) return v;
  return nullptr;
}

std::string env_str(const char* suffix, const char* dflt) {
  const char* v = env_raw(suffix);
  return v ? std::string(v) : std::string(dflt);
}

static bool parse_ll(const char* s, long long* out) {
  if (!s || !*s) return false;
  char* end = nullptr;
  errno = 0;
  long long v = strtoll(s, &end, 0);
  if (errno || end == s) return false;
  // size suffixes: k/m/g (binary)
  if (*end == 'k' || *end == 'K') { v <<= 10; end++; }
  else if (*end == 'm' || *end == 'M') { v <<= 20; end++; }
  else if (*end == 'g' || *end == 'G') { v <<= 30; end++; }
  while (*end == ' ') end++;
  if (*end) return false;
  *out = v;
  return true;
}

long long env_int(const char* suffix, long long dflt) {
  const char* v = env_raw(suffix);
  if (!v) return dflt;
  long long out;
  if (!parse_ll(v, &out)) {
    // The reference unwrap()s here and aborts the process (nthread_…:228-235);
    // we warn and keep the default.
    fprintf(stderr, "[bnet WARN] malformed value '%s' for *_%s, using %lld\n", v, suffix, dflt);
    return dflt;
  }
  return out;
}

const char* env_plain(const char* name) {
  const char* v = getenv(name);
  return (v && *v) ? v : nullptr;
}

long long env_plain_int(const char* name, long long dflt) {
  const char* v = env_plain(name);
  long long out;
  return (v && parse_ll(v, &out)) ? out : dflt;
}

static Config* g_cfg = nullptr;
static std::mutex g_cfg_mu;

static Config parse_config() {
  Config c;
  c.implement = env_str("IMPLEMENT", "BASIC");
  for (auto& ch : c.implement) ch = (char)toupper((unsigned char)ch);
  if (c.implement == "ASYNC" || c.implement == "EPOLL") c.implement = "TOKIO";
  bool async = c.implement == "TOKIO";
  long long ns = env_int("NSTREAMS", 2);
  if (ns < 1) ns = 1;
  if (ns > 64) ns = 64;
  c.nstreams = (int)ns;
  // reference defaults: 1 MiB for BASIC (nthread_…:232-235), 65535 for TOKIO (tokio_…:262-265)
  long long mc = env_int("MIN_CHUNKSIZE", async ? 65535 : 1048576);
  if (mc < 1) mc = 1;
  c.min_chunksize = (size_t)mc;
  long long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  long long aw = env_int("TOKIO_WORKER_THREADS", env_int("ASYNC_WORKERS", ncpu > 4 ? 4 : (ncpu > 0 ? ncpu : 1)));
  if (aw < 1) aw = 1;
  if (aw > 64) aw = 64;
  c.async_workers = (int)aw;
  c.rank = (int)env_plain_int("RANK", -1);
  c.nvl = (int)env_int("NVL", 1);
  c.gdr = (int)env_int("GDR", 1);
  c.wire_compat = (int)env_int("WIRE_COMPAT", 0);
  c.timeout_ms = (int)env_int("TIMEOUT_MS", 0);
  c.spin_us = (int)env_int("SPIN_US", 20);
  c.shm_ring_bytes = (size_t)env_int("SHM_RING_BYTES", 1ll << 20);   // per connection; NCCL opens dozens of them
  c.fault = env_str("FAULT_INJECT", "");
  return c;
}

const Config& Config::get() {
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  if (!g_cfg) g_cfg = new Config(parse_config());
  return *g_cfg;
}

void Config::reload() {
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  // leak the old one on purpose: readers may still hold a reference
  g_cfg = new Config(parse_config());
  g_log_level = -1;
}

// ------------------------------------------------------------------ time / hash
uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}

uint64_t host_hash() {
  static uint64_t h = [] {
    char host[256] = {0};
    gethostname(host, sizeof(host) - 1);
    uint64_t v = fnv1a(host, strlen(host));
    std::ifstream f("/proc/sys/kernel/random/boot_id");
    std::string boot;
    if (f && std::getline(f, boot)) v = fnv1a(boot.data(), boot.size(), v);
    // network namespace identity: abstract unix sockets do not cross it
    char ns[128];
    ssize_t n = readlink("/proc/self/ns/net", ns, sizeof(ns) - 1);
    if (n > 0) v = fnv1a(ns, (size_t)n, v);
    return v ? v : 1;
  }();
  return h;
}

// ---- call breadcrumbs
namespace {
constexpr int kTraceThreads = 128, kTraceDepth = 4;
struct TraceSlot {
  std::atomic<uint32_t> claimed{0};
  std::atomic<int> depth{0};
  std::atomic<uint64_t> tid{0};
  std::atomic<const char*> what[kTraceDepth];
  std::atomic<uint64_t> since[kTraceDepth];
};
TraceSlot g_trace[kTraceThreads];
int trace_slot() {
  thread_local int slot = -2;
  if (slot != -2) return slot;
  slot = -1;
  for (int i = 0; i < kTraceThreads; i++) {
    uint32_t z = 0;
    if (g_trace[i].claimed.compare_exchange_strong(z, 1)) {
      g_trace[i].tid.store((uint64_t)syscall(SYS_gettid), std::memory_order_relaxed);
      slot = i;
      break;
    }
  }
  return slot;   // (slots are never returned: threads that enter the plugin are few and long-lived)
}
}  // namespace

CallScope::CallScope(const char* what) : slot_(trace_slot()), depth_(-1) {
  if (slot_ < 0) return;
  TraceSlot& t = g_trace[slot_];
  int d = t.depth.load(std::memory_order_relaxed);
  if (d < kTraceDepth) {
    t.what[d].store(what, std::memory_order_relaxed);
    t.since[d].store(now_ns(), std::memory_order_relaxed);
    depth_ = d;
  }
  t.depth.store(d + 1, std::memory_order_release);
}
CallScope::~CallScope() {
  if (slot_ < 0) return;
  TraceSlot& t = g_trace[slot_];
  t.depth.store(t.depth.load(std::memory_order_relaxed) - 1, std::memory_order_release);
}
int calltrace_dump(uint64_t older_than_ns) {
  int n = 0;
  uint64_t now = now_ns();
  for (int i = 0; i < kTraceThreads; i++) {
    TraceSlot& t = g_trace[i];
    if (!t.claimed.load(std::memory_order_acquire)) continue;
    int d = t.depth.load(std::memory_order_acquire);
    if (d <= 0) continue;
    if (d > kTraceDepth) d = kTraceDepth;
    uint64_t t0 = t.since[0].load(std::memory_order_relaxed);
    if (t0 >= now || now - t0 < older_than_ns) continue;     // (a scope opened after `now` was read is not old)
    if (t.depth.load(std::memory_order_acquire) <= 0 || t.since[0].load(std::memory_order_relaxed) != t0) continue;   // it moved on
    char line[512];
    int off = snprintf(line, sizeof(line), "[bnet watchdog] pid %d thread %llu inside the plugin:", (int)getpid(),
                       (unsigned long long)t.tid.load(std::memory_order_relaxed));
    for (int k = 0; k < d && off < (int)sizeof(line) - 64; k++) {
      const char* w = t.what[k].load(std::memory_order_relaxed);
      const uint64_t sk = t.since[k].load(std::memory_order_relaxed);
      off += snprintf(line + off, sizeof(line) - off, " %s%s (%.1f ms)", k ? "> " : "", w ? w : "?", sk < now ? (now - sk) / 1e6 : 0.0);
    }
    fprintf(stderr, "%s\n", line);
    n++;
  }
  return n;
}

uint64_t random_u64() {
  static std::mutex mu;
  static std::mt19937_64 rng{std::random_device{}() ^ (uint64_t)getpid() * 0x9E3779B97F4A7C15ull ^ now_ns()};
  std::lock_guard<std::mutex> lk(mu);
  return rng();
}

// ------------------------------------------------------------------ URL / base64
bool parse_user_pass_and_addr(const std::string& raw, UserPassAddr* out) {
  // grammar of the reference regex ^(?:([^:]+):([^@]+)@)?(\S+)$ (utils.rs:181)
  if (raw.empty()) return false;
  for (char ch : raw)
    if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') return false;
  out->user.clear();
  out->pass.clear();
  size_t at = raw.find('@');
  if (at != std::string::npos) {
    size_t colon = raw.find(':');
    if (colon != std::string::npos && colon > 0 && colon + 1 < at && at + 1 < raw.size()) {
      out->user = raw.substr(0, colon);
      out->pass = raw.substr(colon + 1, at - colon - 1);
      out->addr = raw.substr(at + 1);
      return true;
    }
  }
  out->addr = raw;
  return true;
}

std::string base64(const std::string& in) {
  static const char tbl[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string out;
  size_t i = 0;
  while (i + 2 < in.size()) {
    uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8) | (unsigned char)in[i + 2];
    out += tbl[(v >> 18) & 63]; out += tbl[(v >> 12) & 63]; out += tbl[(v >> 6) & 63]; out += tbl[v & 63];
    i += 3;
  }
  if (i + 1 == in.size()) {
    uint32_t v = (unsigned char)in[i] << 16;
    out += tbl[(v >> 18) & 63]; out += tbl[(v >> 12) & 63]; out += "==";
  } else if (i + 2 == in.size()) {
    uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8);
    out += tbl[(v >> 18) & 63]; out += tbl[(v >> 12) & 63]; out += tbl[(v >> 6) & 63]; out += '=';
  }
  return out;
}

// ------------------------------------------------------------------ sockets
socklen_t sockaddr_len(const SockAddr& a) {
  return a.sa.sa_family == AF_INET6 ? sizeof(sockaddr_in6) : sizeof(sockaddr_in);
}

std::string sockaddr_str(const SockAddr& a) {
  char ip[INET6_ADDRSTRLEN] = {0};
  char out[INET6_ADDRSTRLEN + 16];
  if (a.sa.sa_family == AF_INET) {
    inet_ntop(AF_INET, &a.in4.sin_addr, ip, sizeof(ip));
    snprintf(out, sizeof(out), "%s:%u", ip, ntohs(a.in4.sin_port));
  } else if (a.sa.sa_family == AF_INET6) {
    inet_ntop(AF_INET6, &a.in6.sin6_addr, ip, sizeof(ip));
    snprintf(out, sizeof(out), "[%s]:%u", ip, ntohs(a.in6.sin6_port));
  } else {
    snprintf(out, sizeof(out), "<af %d>", a.sa.sa_family);
  }
  return out;
}

bool sockaddr_parse(const std::string& s, SockAddr* out) {
  memset(out, 0, sizeof(*out));
  if (s.empty()) return false;
  if (s[0] == '[') {
    size_t rb = s.find(']');
    if (rb == std::string::npos || rb + 2 > s.size() || s[rb + 1] != ':') return false;
    std::string ip = s.substr(1, rb - 1);
    int port = atoi(s.c_str() + rb + 2);
    out->in6.sin6_family = AF_INET6;
    out->in6.sin6_port = htons((uint16_t)port);
    return inet_pton(AF_INET6, ip.c_str(), &out->in6.sin6_addr) == 1;
  }
  size_t c = s.rfind(':');
  if (c == std::string::npos) return false;
  std::string ip = s.substr(0, c);
  int port = atoi(s.c_str() + c + 1);
  out->in4.sin_family = AF_INET;
  out->in4.sin_port = htons((uint16_t)port);
  return inet_pton(AF_INET, ip.c_str(), &out->in4.sin_addr) == 1;
}

int set_nonblocking(int fd, bool on) {
  int fl = fcntl(fd, F_GETFL, 0);
  if (fl < 0) return -1;
  fl = on ? (fl | O_NONBLOCK) : (fl & ~O_NONBLOCK);
  return fcntl(fd, F_SETFL, fl);
}

int set_nodelay(int fd) {
  int one = 1;
  return setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}

static int wait_fd(int fd, short ev, const std::atomic<bool>* abort, uint64_t deadline_ns) {
  for (;;) {
    if (abort && abort->load(std::memory_order_relaxed)) return kErrRemote;
    int slice = 50;  // ms; bounded so that abort flags are noticed
    if (deadline_ns) {
      uint64_t now = now_ns();
      if (now >= deadline_ns) return kErrTimeout;
      uint64_t left = (deadline_ns - now) / 1000000ull + 1;
      if (left < (uint64_t)slice) slice = (int)left;
    }
    pollfd p{fd, ev, 0};
    int r = poll(&p, 1, slice);
    if (r > 0) return kOk;  // readable/writable or error: let the syscall report it
    if (r < 0 && errno != EINTR) return kErrSystem;
  }
}

int write_all(int fd, const void* buf, size_t n, const std::atomic<bool>* abort, int timeout_ms) {
  const char* p = (const char*)buf;
  uint64_t deadline = timeout_ms > 0 ? now_ns() + (uint64_t)timeout_ms * 1000000ull : 0;
  while (n) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w > 0) {
      p += w;
      n -= (size_t)w;
      if (timeout_ms > 0) deadline = now_ns() + (uint64_t)timeout_ms * 1000000ull;  // progress resets watchdog
      continue;
    }
    if (w == 0) return kErrRemote;
    if (errno == EINTR) continue;
    if (errno == EAGAIN || errno == EWOULDBLOCK) {
      int st = wait_fd(fd, POLLOUT, abort, deadline);
      if (st != kOk) return st;
      continue;
    }
    if (errno == EPIPE || errno == ECONNRESET) return kErrRemote;
    return kErrSystem;
  }
  return kOk;
}

int read_exact(int fd, void* buf, size_t n, const std::atomic<bool>* abort, int timeout_ms) {
  char* p = (char*)buf;
  uint64_t deadline = timeout_ms > 0 ? now_ns() + (uint64_t)timeout_ms * 1000000ull : 0;
  while (n) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r > 0) {
      p += r;
      n -= (size_t)r;
      if (timeout_ms > 0) deadline = now_ns() + (uint64_t)timeout_ms * 1000000ull;
      continue;
    }
    if (r == 0) return kErrRemote;  // EOF before the message was complete
    if (errno == EINTR) continue;
    if (errno == EAGAIN || errno == EWOULDBLOCK) {
      int st = wait_fd(fd, POLLIN, abort, deadline);
      if (st != kOk) return st;
      continue;
    }
    if (errno == ECONNRESET) return kErrRemote;
    return kErrSystem;
  }
  return kOk;
}

// ------------------------------------------------------------------ sysfs
int net_if_speed_mbps(const std::string& ifname) {
  const int kDefault = 10000;
  std::ifstream f("/sys/class/net/" + ifname + "/speed");
  long v = -1;
  if (f && (f >> v) && v > 0) return (int)v;
  return kDefault;
}

std::string net_if_pci_path(const std::string& ifname) {
  char buf[PATH_MAX];
  std::string p = "/sys/class/net/" + ifname + "/device";
  if (realpath(p.c_str(), buf)) return buf;
  return "";
}

}  // namespace bnet
