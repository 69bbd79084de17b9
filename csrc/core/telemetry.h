// Telemetry: per-request spans + Prometheus-style metrics.
//
// Parity with the reference's L7 (reference
// nthread_per_socket_backend.rs:108-212, 529-538, 565-573, 606-618):
//   * root span "BaguaNet-<rank>", child spans isend-<comm>/irecv-<comm>
//     opened in isend/irecv and closed when test() reports completion;
//     exported only when 0 <= RANK <= 7 and a collector address is set
//   * meter "bagua-net": isend_nbytes / irecv_nbytes histograms with
//     boundaries [16,1024,4096,1048576], isend_nbytes_per_second,
//     isend_percentage_of_effective_time, isend_per_second, hold_on_request
//   * a push thread sending the registry to a Pushgateway
//     ([user:pass@]host:port, job BaguaNet, label rank)
// Re-designed: lock-free counters, a bounded span ring, a >=100 ms push period
// (the reference pushes every 200 us), file sinks for boxes without network
// (BNET_TRACE_FILE = Chrome trace JSON, BNET_METRICS_FILE = text exposition).
// Trace wire formats: BAGUA_NET_JAEGER_ADDRESS -> Jaeger collector, Thrift-binary jaeger.thrift Batch POSTed to
// /api/traces (what the reference's opentelemetry-jaeger pipeline sends); BNET_OTLP_ADDRESS -> OTLP/HTTP JSON
// POSTed to /v1/traces.
#pragma once
#include <atomic>
#include <cstdint>
#include <string>

namespace bnet {

struct Histogram {
  static constexpr int kBuckets = 5;  // 4 boundaries + +Inf
  static constexpr uint64_t kBounds[4] = {16, 1024, 4096, 1048576};
  std::atomic<uint64_t> bucket[kBuckets];
  std::atomic<uint64_t> sum{0}, count{0};
  Histogram() { for (auto& b : bucket) b.store(0); }
  void record(uint64_t v) {
    int i = 0;
    while (i < 4 && v > kBounds[i]) i++;
    bucket[i].fetch_add(1, std::memory_order_relaxed);
    sum.fetch_add(v, std::memory_order_relaxed);
    count.fetch_add(1, std::memory_order_relaxed);
  }
};

struct Metrics {
  Histogram isend_nbytes, irecv_nbytes;
  std::atomic<uint64_t> isend_total{0}, irecv_total{0};           // requests
  std::atomic<uint64_t> isend_bytes_total{0}, irecv_bytes_total{0};
  std::atomic<uint64_t> nvl_bytes_total{0}, nvl_kernel_chunks{0}, shm_bytes_total{0}, cma_msgs{0};
  std::atomic<uint64_t> errors_total{0};
  std::atomic<int64_t> hold_on_request{0};                         // in flight
  std::atomic<uint64_t> last_chunk_bytes_per_s{0};                 // isend_nbytes_per_second
  std::atomic<uint64_t> busy_ns{0}, wall_ns{0};                    // -> percentage_of_effective_time
  std::atomic<uint64_t> isend_rate_window_start_ns{0}, isend_rate_window_count{0};
};

enum SpanKind : uint8_t { SPAN_ISEND = 0, SPAN_IRECV = 1, SPAN_IFLUSH = 2, SPAN_ROOT = 3, SPAN_COLL = 4 };

class Telemetry {
 public:
  static Telemetry& get();
  Metrics& m() { return metrics_; }

  bool tracing() const { return tracing_; }
  // returns a span id (0 = tracing off)
  uint64_t span_begin(SpanKind kind, uint64_t comm_id, uint64_t req_id, uint64_t nbytes);
  void span_end(uint64_t span_id, uint64_t nbytes);

  void on_chunk_sent(uint64_t nbytes, uint64_t busy_ns);
  void on_chunk_recv(uint64_t nbytes);

  std::string render_prometheus() const;   // text exposition
  std::string render_trace_json() const;   // Chrome trace-event JSON of finished spans (BNET_TRACE_FILE)
  // collector wire formats; *cursor = how many finished spans the caller has exported already (advanced on return)
  std::string render_jaeger_thrift(uint64_t* cursor) const;   // jaeger.thrift Batch, TBinaryProtocol
  std::string render_otlp_json(uint64_t* cursor) const;       // OTLP/HTTP JSON ExportTraceServiceRequest
  void set_root_attribute(const std::string& socket_devs);    // attribute of the root span (reference nthread_…:132-137)
  std::string span_display_name(int kind, uint64_t comm_id) const;
  int flush();                              // write files / push now; returns #sinks that succeeded
  void shutdown();                          // stop push thread, final flush

  // "host:port" + path + body -> HTTP status (or -1); tiny blocking client
  static int http_send(const std::string& method, const std::string& hostport, const std::string& path,
                       const std::string& content_type, const std::string& body,
                       const std::string& user, const std::string& pass, int timeout_ms);

 private:
  Telemetry();
  ~Telemetry();
  struct Impl;
  Impl* impl_;
  Metrics metrics_;
  bool tracing_ = false;
};

}  // namespace bnet
