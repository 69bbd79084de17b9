"""Telemetry parity tests: spans per request exported as a trace,
Prometheus-style metrics with the reference's instrument names, Pushgateway push with
basic auth, rank gating of the network exporters.  The reference has no tests for any of it."""
import base64
import http.server
import json
import os
import threading

from conftest import run_pair


class _Sink(http.server.BaseHTTPRequestHandler):
    hits = []

    def _take(self):
        n = int(self.headers.get("Content-Length", "0"))
        body = self.rfile.read(n).decode()
        _Sink.hits.append((self.command, self.path, self.headers.get("Authorization"), body))
        self.send_response(200)
        self.send_header("Content-Length", "0")
        self.end_headers()

    do_PUT = do_POST = _take

    def log_message(self, *a):
        pass


def _server():
    _Sink.hits = []
    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), _Sink)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    return srv, srv.server_address[1]


def test_metrics_file_and_trace_file(tmp_path):
    mfile, tfile = tmp_path / "metrics.prom", tmp_path / "trace.json"
    env = {"BNET_NVL": "0", "BNET_METRICS_FILE": str(mfile), "BNET_TRACE_FILE": str(tfile), "RANK": "1",
           "BNET_METRICS_INTERVAL_MS": "100"}
    outs = run_pair(["--sizes", "16,2048,65536,2097152", "--inflight", "4", "--rounds", "2"], env=env)
    assert all(rc == 0 and res["ok"] for rc, res, _ in outs)
    text = mfile.read_text()
    # instrument names and histogram boundaries of the reference (nthread_…:139-182)
    for name in ("isend_nbytes_bucket", "irecv_nbytes_bucket", "isend_nbytes_per_second",
                 "isend_percentage_of_effective_time", "isend_per_second", "hold_on_request"):
        assert name in text, name
    for le in ('le="16"', 'le="1024"', 'le="4096"', 'le="1048576"', 'le="+Inf"'):
        assert le in text
    assert 'handler="all"' in text
    tr = json.loads(tfile.read_text())
    names = {e["name"].split("-")[0] for e in tr["traceEvents"]}
    assert {"isend", "BaguaNet"} <= names or {"irecv", "BaguaNet"} <= names   # both ranks wrote the same file name
    spans = [e for e in tr["traceEvents"] if e["name"].startswith(("isend-", "irecv-"))]
    assert len(spans) == 32 and all(e["dur"] >= 0 and "nbytes" in e["args"] for e in spans)
    assert tr["otherData"]["service"] == "bagua-net"


def test_pushgateway_push_with_basic_auth_and_jaeger_gate():
    srv, port = _server()
    try:
        env = {"BNET_NVL": "0", "RANK": "3", "BAGUA_NET_PROMETHEUS_ADDRESS": f"alice:s3cret@127.0.0.1:{port}",
               "BAGUA_NET_JAEGER_ADDRESS": f"127.0.0.1:{port}", "BNET_METRICS_INTERVAL_MS": "100"}
        outs = run_pair(["--sizes", "4096,1048577", "--inflight", "4", "--rounds", "2"], env=env)
        assert all(rc == 0 and res["ok"] for rc, res, _ in outs)
    finally:
        srv.shutdown()
    puts = [h for h in _Sink.hits if h[0] == "PUT"]
    assert puts, _Sink.hits
    assert all(p[1] == "/metrics/job/BaguaNet/rank/3" for p in puts)       # job + rank label like the reference
    assert puts[0][2] == "Basic " + base64.b64encode(b"alice:s3cret").decode()
    assert "isend_nbytes_bucket" in puts[-1][3]
    posts = [h for h in _Sink.hits if h[0] == "POST" and h[1] == "/api/traces"]
    assert posts and "traceEvents" in posts[-1][3]

    # ranks outside 0..7 do not export traces (reference gate: nthread_…:109-111)
    srv, port = _server()
    try:
        env = {"BNET_NVL": "0", "RANK": "9", "BAGUA_NET_JAEGER_ADDRESS": f"127.0.0.1:{port}"}
        outs = run_pair(["--sizes", "4096", "--inflight", "2", "--rounds", "1"], env=env)
        assert all(rc == 0 and res["ok"] for rc, res, _ in outs)
    finally:
        srv.shutdown()
    assert not [h for h in _Sink.hits if h[0] == "POST"]


def test_metrics_counters_in_process():
    import numpy as np

    from bagua_net_b200 import utils
    from bagua_net_b200.utils.abi import NetPlugin

    os.environ["BNET_NVL"] = "1"
    utils.reload_config()
    p = NetPlugin(8)
    p.init()
    h, l = p.listen(0)
    s = p.connect(h)
    r = p.accept(l)
    src, dst = np.arange(5000, dtype=np.uint8), np.zeros(5000, dtype=np.uint8)
    rq, sq = p.irecv(r, dst.ctypes.data, 5000), p.isend(s, src.ctypes.data, 5000)
    assert p.wait(sq) == 5000 and p.wait(rq) == 5000 and (src == dst).all()
    text = utils.metrics_text()
    assert "bnet_isend_requests_total" in text and "bnet_shm_bytes_total" in text
    line = [ln for ln in text.splitlines() if ln.startswith("bnet_isend_requests_total")][0]
    assert float(line.split()[-1]) >= 1
    p.close_send(s), p.close_recv(r), p.close_listen(l)
    os.environ.pop("BNET_NVL")
    utils.reload_config()
