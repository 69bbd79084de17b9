#include "core/telemetry.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <poll.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <sstream>
#include <thread>
#include <vector>

#include "core/common.h"

namespace bnet {

constexpr uint64_t Histogram::kBounds[4];

namespace {
struct Span {
  uint64_t id;
  uint64_t comm_id, req_id, nbytes;
  uint64_t t0, t1;
  SpanKind kind;
  uint32_t tid;
};
constexpr size_t kOpenSlots = 4096;   // open spans (power of two)
constexpr size_t kDoneRing = 65536;   // finished spans kept for export
}  // namespace

struct Telemetry::Impl {
  std::string jaeger_addr, prom_addr, trace_file, metrics_file;
  UserPassAddr prom;
  int rank = -1;
  int interval_ms = 1000;
  uint64_t t_start = 0;
  // span storage
  std::mutex span_mu;
  std::vector<Span> open;     // indexed by id % kOpenSlots
  std::vector<Span> done;     // ring
  size_t done_head = 0, done_count = 0;
  uint64_t done_total = 0;            // spans ever finished (export cursors are positions in this sequence)
  uint64_t sent_jaeger = 0, sent_otlp = 0;
  std::atomic<uint64_t> next_id{1};
  uint64_t root_id = 0;
  std::string otlp_addr;
  uint64_t trace_hi = 0, trace_lo = 0;   // 128-bit trace id of this plugin instance (one trace per process, like the
  uint64_t span_salt = 0;                // reference's root span per BaguaNet instance, nthread_…:132-137)
  int64_t epoch_off_ns = 0;              // realtime - now_ns() at start: span timestamps are exported in Unix time
  std::string root_devs;                 // attribute "socket_devs" of the root span
  // push thread
  std::thread pusher;
  std::mutex mu;
  std::condition_variable cv;
  bool stop = false;
};

Telemetry& Telemetry::get() {
  // intentionally leaked (workers may still record while the process exits); the final
  // flush — root span end + last push, the reference does it in Drop (nthread_…:652-658) —
  // runs from atexit instead of a destructor.
  static Telemetry* t = [] {
    Telemetry* x = new Telemetry();
    atexit([] { Telemetry::get().shutdown(); });
    return x;
  }();
  return *t;
}

Telemetry::Telemetry() : impl_(new Impl) {
  Impl& I = *impl_;
  const Config& cfg = Config::get();
  I.rank = cfg.rank;
  I.t_start = now_ns();
  I.jaeger_addr = env_str("JAEGER_ADDRESS", "");
  I.otlp_addr = env_str("OTLP_ADDRESS", "");
  {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    I.epoch_off_ns = (int64_t)ts.tv_sec * 1000000000ll + ts.tv_nsec - (int64_t)I.t_start;
    I.trace_hi = random_u64();
    I.trace_lo = random_u64() | 1;
    I.span_salt = random_u64() & ~0xffffffffull;   // span ids: salt | per-process counter (never zero)
  }
  I.prom_addr = env_str("PROMETHEUS_ADDRESS", "");
  I.trace_file = env_str("TRACE_FILE", "");
  I.metrics_file = env_str("METRICS_FILE", "");
  I.interval_ms = (int)env_int("METRICS_INTERVAL_MS", 1000);
  if (I.interval_ms < 100) I.interval_ms = 100;
  // reference gate: exporters only on ranks 0..7 (nthread_…:109-111); file sinks are always allowed
  bool rank_ok = I.rank >= 0 && I.rank <= 7;
  if (!rank_ok) I.jaeger_addr.clear();
  if (!rank_ok) I.otlp_addr.clear();
  tracing_ = !I.jaeger_addr.empty() || !I.otlp_addr.empty() || !I.trace_file.empty();
  if (tracing_) {
    I.open.resize(kOpenSlots);
    I.done.resize(kDoneRing);
    I.root_id = span_begin(SPAN_ROOT, 0, 0, 0);
  }
  if (!I.prom_addr.empty() && !parse_user_pass_and_addr(I.prom_addr, &I.prom)) I.prom_addr.clear();
  if (!I.prom_addr.empty() || !I.metrics_file.empty() || !I.jaeger_addr.empty() || !I.otlp_addr.empty()) {
    I.pusher = std::thread([this] {
      Impl& I = *impl_;
      std::unique_lock<std::mutex> lk(I.mu);
      while (!I.stop) {
        I.cv.wait_for(lk, std::chrono::milliseconds(I.interval_ms));
        if (I.stop) break;
        lk.unlock();
        flush();
        lk.lock();
      }
    });
  }
}

Telemetry::~Telemetry() { shutdown(); delete impl_; }

void Telemetry::shutdown() {
  Impl& I = *impl_;
  {
    std::lock_guard<std::mutex> lk(I.mu);
    if (I.stop) return;
    I.stop = true;
  }
  I.cv.notify_all();
  if (I.pusher.joinable()) I.pusher.join();
  if (tracing_ && I.root_id) span_end(I.root_id, 0);
  flush();
}

uint64_t Telemetry::span_begin(SpanKind kind, uint64_t comm_id, uint64_t req_id, uint64_t nbytes) {
  if (!tracing_) return 0;
  Impl& I = *impl_;
  uint64_t id = I.next_id.fetch_add(1, std::memory_order_relaxed);
  Span s{id, comm_id, req_id, nbytes, now_ns(), 0, kind, (uint32_t)(uintptr_t)pthread_self()};
  std::lock_guard<std::mutex> lk(I.span_mu);
  I.open[id % kOpenSlots] = s;  // a slot collision only loses an ancient unfinished span
  return id;
}

void Telemetry::span_end(uint64_t span_id, uint64_t nbytes) {
  if (!tracing_ || !span_id) return;
  Impl& I = *impl_;
  std::lock_guard<std::mutex> lk(I.span_mu);
  Span& s = I.open[span_id % kOpenSlots];
  if (s.id != span_id) return;
  s.t1 = now_ns();
  if (nbytes) s.nbytes = nbytes;
  I.done[(I.done_head + I.done_count) % kDoneRing] = s;
  if (I.done_count < kDoneRing) I.done_count++;
  else I.done_head = (I.done_head + 1) % kDoneRing;
  I.done_total++;
  s.id = 0;
}

void Telemetry::on_chunk_sent(uint64_t nbytes, uint64_t busy_ns) {
  Metrics& M = metrics_;
  M.isend_nbytes.record(nbytes);
  M.isend_bytes_total.fetch_add(nbytes, std::memory_order_relaxed);
  M.busy_ns.fetch_add(busy_ns, std::memory_order_relaxed);
  if (busy_ns) M.last_chunk_bytes_per_s.store(nbytes * 1000000000ull / busy_ns, std::memory_order_relaxed);
}

void Telemetry::on_chunk_recv(uint64_t nbytes) {
  metrics_.irecv_nbytes.record(nbytes);
  metrics_.irecv_bytes_total.fetch_add(nbytes, std::memory_order_relaxed);
}

static void render_hist(std::ostringstream& o, const char* name, const Histogram& h, const std::string& lbl) {
  o << "# TYPE " << name << " histogram\n";
  uint64_t cum = 0;
  for (int i = 0; i < 4; i++) {
    cum += h.bucket[i].load();
    o << name << "_bucket{" << lbl << ",le=\"" << Histogram::kBounds[i] << "\"} " << cum << "\n";
  }
  cum += h.bucket[4].load();
  o << name << "_bucket{" << lbl << ",le=\"+Inf\"} " << cum << "\n";
  o << name << "_sum{" << lbl << "} " << h.sum.load() << "\n";
  o << name << "_count{" << lbl << "} " << h.count.load() << "\n";
}

std::string Telemetry::render_prometheus() const {
  const Metrics& M = metrics_;
  const Impl& I = *impl_;
  std::ostringstream o;
  std::string lbl = "handler=\"all\"";  // reference label (nthread_…:23-25)
  render_hist(o, "isend_nbytes", M.isend_nbytes, lbl);
  render_hist(o, "irecv_nbytes", M.irecv_nbytes, lbl);
  uint64_t wall = now_ns() - I.t_start;
  double eff = wall ? 100.0 * (double)M.busy_ns.load() / (double)wall : 0.0;
  auto gauge = [&](const char* n, double v) {
    o << "# TYPE " << n << " gauge\n" << n << "{" << lbl << "} " << v << "\n";
  };
  auto counter = [&](const char* n, uint64_t v) {
    o << "# TYPE " << n << " counter\n" << n << "{" << lbl << "} " << v << "\n";
  };
  gauge("isend_nbytes_per_second", (double)M.last_chunk_bytes_per_s.load());
  gauge("isend_percentage_of_effective_time", eff);
  gauge("isend_per_second", wall ? (double)M.isend_total.load() * 1e9 / (double)wall : 0.0);
  gauge("hold_on_request", (double)M.hold_on_request.load());
  counter("bnet_isend_requests_total", M.isend_total.load());
  counter("bnet_irecv_requests_total", M.irecv_total.load());
  counter("bnet_isend_bytes_total", M.isend_bytes_total.load());
  counter("bnet_irecv_bytes_total", M.irecv_bytes_total.load());
  counter("bnet_nvl_bytes_total", M.nvl_bytes_total.load());
  counter("bnet_nvl_kernel_chunks_total", M.nvl_kernel_chunks.load());
  counter("bnet_shm_bytes_total", M.shm_bytes_total.load());
  counter("bnet_cma_messages_total", M.cma_msgs.load());
  counter("bnet_errors_total", M.errors_total.load());
  return o.str();
}

static const char* span_name(SpanKind k) {
  switch (k) {
    case SPAN_ISEND: return "isend";
    case SPAN_IRECV: return "irecv";
    case SPAN_IFLUSH: return "iflush";
    case SPAN_ROOT: return "BaguaNet";
    default: return "coll";
  }
}

std::string Telemetry::render_trace_json() const {
  Impl& I = *impl_;
  std::ostringstream o;
  o << "{\"traceEvents\":[";
  std::lock_guard<std::mutex> lk(I.span_mu);
  bool first = true;
  for (size_t i = 0; i < I.done_count; i++) {
    const Span& s = I.done[(I.done_head + i) % kDoneRing];
    if (!first) o << ",";
    first = false;
    o << "{\"name\":\"" << span_name(s.kind) << "-" << (s.kind == SPAN_ROOT ? (uint64_t)(I.rank < 0 ? 0 : I.rank) : s.comm_id)
      << "\",\"ph\":\"X\",\"pid\":" << (I.rank < 0 ? 0 : I.rank) << ",\"tid\":" << s.tid
      << ",\"ts\":" << (double)(s.t0 - I.t_start) / 1000.0 << ",\"dur\":" << (double)(s.t1 - s.t0) / 1000.0
      << ",\"args\":{\"id\":" << s.req_id << ",\"nbytes\":" << s.nbytes << "}}";
  }
  o << "],\"displayTimeUnit\":\"ns\",\"otherData\":{\"service\":\"bagua-net\",\"rank\":" << I.rank << "}}";
  return o.str();
}


// ---- trace export in the collectors' own wire formats -------------------------------------------------------
// The reference ships its spans through opentelemetry-jaeger's collector pipeline: HTTP POST of a Thrift-binary
// `jaeger.thrift` Batch to http://<addr>/api/traces (reference nthread_…:113-130).  That is what
// BAGUA_NET_JAEGER_ADDRESS speaks here as well (hand-encoded TBinaryProtocol, no Thrift library needed);
// BNET_OTLP_ADDRESS additionally speaks OTLP/HTTP JSON (POST /v1/traces), which current Jaeger and every
// OpenTelemetry collector ingest.  Both export only the spans finished since the previous export.
namespace {
struct ThriftOut {
  std::string b;
  void u8(uint8_t v) { b.push_back((char)v); }
  void i16(uint16_t v) { u8(v >> 8); u8(v & 0xff); }
  void i32(uint32_t v) { for (int s = 24; s >= 0; s -= 8) u8((v >> s) & 0xff); }
  void i64(uint64_t v) { for (int s = 56; s >= 0; s -= 8) u8((v >> s) & 0xff); }
  void field(uint8_t type, uint16_t id) { u8(type); i16(id); }
  void str(const std::string& v) { i32((uint32_t)v.size()); b += v; }
  void stop() { u8(0); }
};
enum : uint8_t { T_BOOL = 2, T_DOUBLE = 4, T_I32 = 8, T_I64 = 10, T_STRING = 11, T_STRUCT = 12, T_LIST = 15 };
void tag_str(ThriftOut& o, const char* k, const std::string& v) {
  o.field(T_STRING, 1); o.str(k);
  o.field(T_I32, 2); o.i32(0);        // TagType STRING
  o.field(T_STRING, 3); o.str(v);
  o.stop();
}
void tag_long(ThriftOut& o, const char* k, int64_t v) {
  o.field(T_STRING, 1); o.str(k);
  o.field(T_I32, 2); o.i32(3);        // TagType LONG
  o.field(T_I64, 6); o.i64((uint64_t)v);
  o.stop();
}
std::string hex(uint64_t v) {
  char b[17];
  snprintf(b, sizeof(b), "%016llx", (unsigned long long)v);
  return b;
}
}  // namespace

std::string Telemetry::span_display_name(int kind, uint64_t comm_id) const {
  const Impl& I = *impl_;
  std::string n = span_name((SpanKind)kind);
  n += "-";
  n += std::to_string(kind == SPAN_ROOT ? (uint64_t)(I.rank < 0 ? 0 : I.rank) : comm_id);
  return n;
}

// jaeger.thrift Batch{1: Process{1: serviceName, 2: tags}, 2: list<Span>} of the spans finished after *cursor
std::string Telemetry::render_jaeger_thrift(uint64_t* cursor) const {
  Impl& I = *impl_;
  std::lock_guard<std::mutex> lk(I.span_mu);
  uint64_t first_kept = I.done_total - I.done_count;
  uint64_t from = *cursor < first_kept ? first_kept : *cursor;
  uint32_t n = (uint32_t)(I.done_total - from);
  ThriftOut o;
  o.field(T_STRUCT, 1);
  o.field(T_STRING, 1); o.str("bagua-net");
  o.field(T_LIST, 2); o.u8(T_STRUCT); o.i32(2);
  tag_long(o, "rank", I.rank);
  tag_str(o, "plugin", "bnet");
  o.stop();
  o.field(T_LIST, 2); o.u8(T_STRUCT); o.i32(n);
  const uint64_t root_span = I.span_salt | (I.root_id & 0xffffffffull);
  for (uint64_t k = from; k < I.done_total; k++) {
    const Span& s = I.done[(I.done_head + (k - first_kept)) % kDoneRing];
    o.field(T_I64, 1); o.i64(I.trace_lo);
    o.field(T_I64, 2); o.i64(I.trace_hi);
    o.field(T_I64, 3); o.i64(I.span_salt | (s.id & 0xffffffffull));
    o.field(T_I64, 4); o.i64(s.kind == SPAN_ROOT ? 0 : root_span);
    o.field(T_STRING, 5); o.str(span_display_name(s.kind, s.comm_id));
    o.field(T_I32, 7); o.i32(1);   // sampled
    o.field(T_I64, 8); o.i64((uint64_t)(((int64_t)s.t0 + I.epoch_off_ns) / 1000));
    o.field(T_I64, 9); o.i64((s.t1 - s.t0) / 1000);
    if (s.kind == SPAN_ROOT) {
      o.field(T_LIST, 10); o.u8(T_STRUCT); o.i32(1);
      tag_str(o, "socket_devs", I.root_devs);
    } else {
      o.field(T_LIST, 10); o.u8(T_STRUCT); o.i32(2);
      tag_long(o, "id", (int64_t)s.req_id);
      tag_long(o, "nbytes", (int64_t)s.nbytes);
    }
    o.stop();
  }
  o.stop();
  *cursor = I.done_total;
  return n ? o.b : std::string();
}

// OTLP/HTTP JSON ExportTraceServiceRequest of the spans finished after *cursor
std::string Telemetry::render_otlp_json(uint64_t* cursor) const {
  Impl& I = *impl_;
  std::lock_guard<std::mutex> lk(I.span_mu);
  uint64_t first_kept = I.done_total - I.done_count;
  uint64_t from = *cursor < first_kept ? first_kept : *cursor;
  if (from == I.done_total) { *cursor = I.done_total; return std::string(); }
  const std::string trace_id = hex(I.trace_hi) + hex(I.trace_lo);
  const uint64_t root_span = I.span_salt | (I.root_id & 0xffffffffull);
  std::ostringstream o;
  o << "{\"resourceSpans\":[{\"resource\":{\"attributes\":[{\"key\":\"service.name\",\"value\":{\"stringValue\":\"bagua-net\"}},"
    << "{\"key\":\"rank\",\"value\":{\"intValue\":\"" << I.rank << "\"}}]},\"scopeSpans\":[{\"scope\":{\"name\":\"bnet\"},\"spans\":[";
  bool first = true;
  for (uint64_t k = from; k < I.done_total; k++) {
    const Span& s = I.done[(I.done_head + (k - first_kept)) % kDoneRing];
    if (!first) o << ",";
    first = false;
    o << "{\"traceId\":\"" << trace_id << "\",\"spanId\":\"" << hex(I.span_salt | (s.id & 0xffffffffull)) << "\"";
    if (s.kind != SPAN_ROOT) o << ",\"parentSpanId\":\"" << hex(root_span) << "\"";
    o << ",\"name\":\"" << span_display_name(s.kind, s.comm_id) << "\",\"kind\":" << (s.kind == SPAN_ISEND ? 4 : s.kind == SPAN_IRECV ? 5 : 1)
      << ",\"startTimeUnixNano\":\"" << (int64_t)s.t0 + I.epoch_off_ns << "\",\"endTimeUnixNano\":\"" << (int64_t)s.t1 + I.epoch_off_ns
      << "\",\"attributes\":[{\"key\":\"id\",\"value\":{\"intValue\":\"" << s.req_id << "\"}},{\"key\":\"nbytes\",\"value\":{\"intValue\":\""
      << s.nbytes << "\"}}]}";
  }
  o << "]}]}]}";
  *cursor = I.done_total;
  return o.str();
}

void Telemetry::set_root_attribute(const std::string& socket_devs) { impl_->root_devs = socket_devs; }

int Telemetry::http_send(const std::string& method, const std::string& hostport, const std::string& path,
                         const std::string& ctype, const std::string& body, const std::string& user,
                         const std::string& pass, int timeout_ms) {
  size_t c = hostport.rfind(':');
  if (c == std::string::npos) return -1;
  std::string host = hostport.substr(0, c), port = hostport.substr(c + 1);
  if (!host.empty() && host[0] == '[' && host.back() == ']') host = host.substr(1, host.size() - 2);
  addrinfo hints{}, *res = nullptr;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || !res) return -1;
  int fd = socket(res->ai_family, SOCK_STREAM, 0);
  int status = -1;
  if (fd >= 0) {
    timeval tv{timeout_ms / 1000, (timeout_ms % 1000) * 1000};
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
      std::ostringstream req;
      req << method << " " << path << " HTTP/1.1\r\nHost: " << hostport << "\r\nContent-Type: " << ctype
          << "\r\nContent-Length: " << body.size() << "\r\nConnection: close\r\n";
      if (!user.empty()) req << "Authorization: Basic " << base64(user + ":" + pass) << "\r\n";
      req << "\r\n" << body;
      std::string r = req.str();
      if (write_all(fd, r.data(), r.size(), nullptr, timeout_ms) == kOk) {
        char buf[64] = {0};
        ssize_t n = recv(fd, buf, sizeof(buf) - 1, 0);
        if (n > 12 && !strncmp(buf, "HTTP/", 5)) status = atoi(buf + 9);
      }
    }
    close(fd);
  }
  freeaddrinfo(res);
  return status;
}

int Telemetry::flush() {
  Impl& I = *impl_;
  int ok = 0;
  if (!I.metrics_file.empty() || !I.prom_addr.empty()) {
    std::string text = render_prometheus();
    if (!I.metrics_file.empty()) {
      std::string tmp = I.metrics_file + ".tmp";
      FILE* f = fopen(tmp.c_str(), "w");
      if (f) {
        fwrite(text.data(), 1, text.size(), f);
        fclose(f);
        if (rename(tmp.c_str(), I.metrics_file.c_str()) == 0) ok++;
      }
    }
    if (!I.prom_addr.empty()) {
      char path[128];
      snprintf(path, sizeof(path), "/metrics/job/BaguaNet/rank/%d", I.rank);
      int st = http_send("PUT", I.prom.addr, path, "text/plain; version=0.0.4", text, I.prom.user, I.prom.pass, 500);
      if (st >= 200 && st < 300) ok++;
    }
  }
  if (tracing_) {
    if (!I.trace_file.empty()) {
      std::string js = render_trace_json();
      FILE* f = fopen(I.trace_file.c_str(), "w");
      if (f) {
        fwrite(js.data(), 1, js.size(), f);
        fclose(f);
        ok++;
      }
    }
    if (!I.jaeger_addr.empty()) {
      uint64_t cur = I.sent_jaeger;
      std::string batch = render_jaeger_thrift(&cur);
      if (batch.empty()) {
        I.sent_jaeger = cur;
      } else {
        int st = http_send("POST", I.jaeger_addr, "/api/traces?format=jaeger.thrift", "application/x-thrift", batch, "", "", 500);
        if (st >= 200 && st < 300) { ok++; I.sent_jaeger = cur; }   // not accepted: the same spans go out again next time
      }
    }
    if (!I.otlp_addr.empty()) {
      uint64_t cur = I.sent_otlp;
      std::string body = render_otlp_json(&cur);
      if (body.empty()) {
        I.sent_otlp = cur;
      } else {
        int st = http_send("POST", I.otlp_addr, "/v1/traces", "application/json", body, "", "", 500);
        if (st >= 200 && st < 300) { ok++; I.sent_otlp = cur; }
      }
    }
  }
  return ok;
}

}  // namespace bnet
