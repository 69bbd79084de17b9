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
        raw = self.rfile.read(n)
        body = raw if self.headers.get("Content-Type") == "application/x-thrift" else raw.decode()
        _Sink.hits.append((self.command, self.path, self.headers.get("Authorization"), body, self.headers.get("Content-Type")))
        self.send_response(200)
        self.send_header("Content-Length", "0")
        self.end_headers()

    do_PUT = do_POST = _take

    def log_message(self, *a):
        pass


# ---- a minimal TBinaryProtocol reader: enough to check a jaeger.thrift Batch against its schema -------------
import struct  # noqa: E402

T_BOOL, T_BYTE, T_DOUBLE, T_I16, T_I32, T_I64, T_STRING, T_STRUCT, T_LIST = 2, 3, 4, 6, 8, 10, 11, 12, 15


class _Thrift:
    def __init__(self, data: bytes):
        self.d, self.p = data, 0

    def take(self, fmt):
        v = struct.unpack_from(">" + fmt, self.d, self.p)
        self.p += struct.calcsize(">" + fmt)
        return v[0]

    def value(self, t):
        if t == T_BOOL or t == T_BYTE:
            return self.take("b")
        if t == T_DOUBLE:
            return self.take("d")
        if t == T_I16:
            return self.take("h")
        if t == T_I32:
            return self.take("i")
        if t == T_I64:
            return self.take("q")
        if t == T_STRING:
            n = self.take("i")
            v = self.d[self.p:self.p + n]
            self.p += n
            return v.decode()
        if t == T_STRUCT:
            return self.struct()
        if t == T_LIST:
            et, n = self.take("b"), self.take("i")
            return [self.value(et) for _ in range(n)]
        raise AssertionError(f"unexpected thrift type {t}")

    def struct(self):
        out = {}
        while True:
            t = self.take("b")
            if t == 0:
                return out
            fid = self.take("h")
            out[fid] = (t, self.value(t))


def _check_jaeger_batch(raw: bytes):
    """jaeger.thrift: Batch{1: Process{1: serviceName, 2: list<Tag>}, 2: list<Span>};
    Span{1,2: traceId low/high i64, 3: spanId, 4: parentSpanId, 5: operationName, 7: flags i32, 8: startTime us, 9: duration us,
    10: list<Tag>}; Tag{1: key, 2: vType i32, 3: vStr | 6: vLong}."""
    rd = _Thrift(raw)
    batch = rd.struct()
    assert rd.p == len(raw), "trailing bytes after the Batch"
    assert batch[1][0] == T_STRUCT and batch[2][0] == T_LIST
    proc = batch[1][1]
    assert proc[1] == (T_STRING, "bagua-net")
    spans = batch[2][1]
    assert spans
    names = []
    for sp in spans:
        for fid, ty in ((1, T_I64), (2, T_I64), (3, T_I64), (4, T_I64), (5, T_STRING), (7, T_I32), (8, T_I64), (9, T_I64)):
            assert sp[fid][0] == ty, (fid, sp[fid])
        assert sp[3][1] != 0 and sp[9][1] >= 0
        assert sp[8][1] > 1_600_000_000_000_000, "startTime must be microseconds since the Unix epoch"
        tags = {t[1][1]: t for t in sp[10][1]}
        for t in tags.values():
            assert t[2][0] == T_I32 and ((t[2][1] == 0 and 3 in t) or (t[2][1] == 3 and 6 in t))
        names.append((sp[5][1], sp[3][1]))
        if not sp[5][1].startswith("BaguaNet-"):
            assert {"id", "nbytes"} <= set(tags) and sp[4][1] != 0     # children point at the root span
    return names


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
    # the Jaeger collector endpoint gets a Thrift-binary jaeger.thrift Batch, like the reference's
    # opentelemetry-jaeger pipeline sends (nthread_…:113-130)
    posts = [h for h in _Sink.hits if h[0] == "POST" and h[1].startswith("/api/traces")]
    assert posts and all(h[4] == "application/x-thrift" for h in posts)
    pairs = [n for h in posts for n in _check_jaeger_batch(h[3])]
    names = [n for n, _ in pairs]
    kids = [pr for pr in pairs if pr[0].startswith(("isend-", "irecv-"))]
    assert len(kids) == 32 and len(set(kids)) == 32       # 16 sends + 16 receives, every span exported exactly once
    assert any(n == "BaguaNet-3" for n in names)          # root span, exported at shutdown

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


def test_otlp_http_json_export():
    """BNET_OTLP_ADDRESS: OTLP/HTTP JSON to /v1/traces (what OpenTelemetry collectors and current Jaeger ingest)."""
    srv, port = _server()
    try:
        env = {"BNET_NVL": "0", "RANK": "0", "BNET_OTLP_ADDRESS": f"127.0.0.1:{port}", "BNET_METRICS_INTERVAL_MS": "100"}
        outs = run_pair(["--sizes", "4096,70000", "--inflight", "2", "--rounds", "2"], env=env)
        assert all(rc == 0 and res["ok"] for rc, res, _ in outs)
    finally:
        srv.shutdown()
    posts = [h for h in _Sink.hits if h[0] == "POST" and h[1] == "/v1/traces"]
    assert posts
    spans = []
    for h in posts:
        doc = json.loads(h[3])
        for rs in doc["resourceSpans"]:
            attrs = {a["key"]: a["value"] for a in rs["resource"]["attributes"]}
            assert attrs["service.name"] == {"stringValue": "bagua-net"}
            for ss in rs["scopeSpans"]:
                spans += ss["spans"]
    assert spans
    for sp in spans:
        assert len(sp["traceId"]) == 32 and len(sp["spanId"]) == 16 and int(sp["traceId"], 16) and int(sp["spanId"], 16)
        assert int(sp["endTimeUnixNano"]) >= int(sp["startTimeUnixNano"]) > 1_600_000_000_000_000_000
        assert {a["key"] for a in sp["attributes"]} >= {"id", "nbytes"}
    kids = [sp for sp in spans if sp["name"].startswith(("isend-", "irecv-"))]
    assert len(kids) == 16 and all(len(sp["parentSpanId"]) == 16 for sp in kids)     # 8 sends + 8 receives
    assert len({(sp["traceId"], sp["spanId"]) for sp in kids}) == 16                   # every span exactly once


def test_collectives_over_the_mesh_are_spans_around_their_messages(tmp_path):
    """An all-reduce over the mesh of plugin connections opens a "coll-<rank>" span (its bytes as an attribute); the isend /
    irecv spans of its messages fall inside it in the exported trace."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world, count = 3, 40000
    procs = []
    for r in range(world):
        env = dict(os.environ, BNET_FAKE_CUDA="1", BNET_NVL="1", RANK=str(r), BNET_TRACE_FILE=str(tmp_path / f"trace{r}.json"),
                   PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        procs.append(subprocess.Popen([sys.executable, os.path.join(root, "tests", "tmesh_worker.py"), str(r), str(world), str(tmp_path),
                                       str(count), "f32", "f32", "16384", "4", "2", "two-shot", "0"], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        assert json.loads([ln for ln in o.splitlines() if ln.startswith("{")][-1])["ok"]
    for r in range(world):
        tr = json.loads((tmp_path / f"trace{r}.json").read_text())
        colls = [e for e in tr["traceEvents"] if e["name"] == f"coll-{r}"]
        assert len(colls) == 2 and all(e["args"]["nbytes"] == count * 4 for e in colls), colls        # two rounds
        msgs = [e for e in tr["traceEvents"] if e["name"].startswith(("isend-", "irecv-")) and e["args"]["nbytes"] > 4]
        assert msgs
        lo, hi = min(c["ts"] for c in colls), max(c["ts"] + c["dur"] for c in colls)
        assert all(lo <= e["ts"] and e["ts"] + e["dur"] <= hi + 1 for e in msgs), "a message span lies outside the collectives"
