"""Prometheus metrics (reference: pico-core misc/observability/metrics/Metrics.h and the
``server`` daemon flags --enable_metrics --metrics_ip --metrics_port(8001) --metrics_url,
openembedding/entry/server.cc:7-12,35-36).

Series kept name-compatible: ``ps_wait_duration_ms`` / ``ps_wait_request_count`` (labelled by
storage, handler; pico-ps handler/Handler.cpp:14-18,49-56) and ``ps_request_duration_ms`` /
``ps_requests_total`` / ``ps_errors_total`` (labelled by request_type; service/Service.cpp:937-974).
"""
import threading
import time

try:
    import prometheus_client as _pc
except Exception:      # pragma: no cover - prometheus_client is in the image
    _pc = None

_registry = None
_metrics = {}
_lock = threading.Lock()
_server_started = False


def _get(name, kind, doc, labels):
    global _registry
    if _pc is None:
        return None
    with _lock:
        if _registry is None:
            _registry = _pc.CollectorRegistry()
        if name not in _metrics:
            cls = {"counter": _pc.Counter, "gauge": _pc.Gauge, "histogram": _pc.Histogram}[kind]
            kw = {"buckets": (0.01, 0.05, 0.1, 0.5, 1, 5, 10, 50, 100, 500, 1000, 5000)} if kind == "histogram" else {}
            _metrics[name] = cls(name, doc, labels, registry=_registry, **kw)
        return _metrics[name]


def observe_wait(storage, handler, ms):
    h = _get("ps_wait_duration_ms", "histogram", "client wait per handler call (ms)", ["storage", "handler"])
    c = _get("ps_wait_request_count", "counter", "client handler calls", ["storage", "handler"])
    if h is not None:
        h.labels(str(storage), handler).observe(ms)
        c.labels(str(storage), handler).inc()


def observe_request(request_type, ms, error=False):
    h = _get("ps_request_duration_ms", "histogram", "request latency (ms)", ["request_type"])
    c = _get("ps_requests_total", "counter", "requests", ["request_type"])
    e = _get("ps_errors_total", "counter", "failed requests", ["request_type"])
    if h is not None:
        h.labels(request_type).observe(ms)
        c.labels(request_type).inc()
        if error:
            e.labels(request_type).inc()


def set_gauge(name, value, doc="", **labels):
    g = _get(name, "gauge", doc or name, sorted(labels))
    if g is not None:
        (g.labels(**labels) if labels else g).set(value)


class timed_request:
    def __init__(self, request_type):
        self.rt = request_type

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, et, ev, tb):
        observe_request(self.rt, (time.perf_counter() - self.t0) * 1e3, error=et is not None)
        return False


def start_exposer(ip="0.0.0.0", port=8001):
    """HTTP /metrics endpoint (prometheus exposer of the server daemon)."""
    global _server_started
    if _pc is None:
        raise RuntimeError("prometheus_client not available")
    _get("exb_up", "gauge", "process is up", [])
    _metrics["exb_up"].set(1)
    if not _server_started:
        _pc.start_http_server(port, addr=ip, registry=_registry)
        _server_started = True
    return "%s:%d" % (ip, port)


def render():
    """current exposition text (tests / controller introspection)"""
    if _pc is None or _registry is None:
        return ""
    return _pc.generate_latest(_registry).decode()
