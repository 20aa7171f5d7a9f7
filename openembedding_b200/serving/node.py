"""Serving node: a process that hosts read-only shards of exported models.

Reference: the ``server`` daemon (openembedding/entry/server.cc) + read-only pulls
(``read_only_pull`` handler: pick_one_replica, get_weights, never inserts, missing rows get
the initializer value -- openembedding/server/EmbeddingPullOperator.cpp:50,179-181) + the
peer-to-peer restore of a replaced node (EmbeddingRestoreOperator.cpp:19-152).

Transport is plain HTTP (stdlib): serving traffic is request/response with small payloads
and lives outside the NVLink training domain; row payloads are raw little-endian bytes.
"""
import ctypes
import json
import os
import threading
import zlib
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from urllib.parse import parse_qs, urlparse

import numpy as np

from .. import _native
from ..checkpoint import iter_shard_file, read_model_meta
from ..config import DTYPES, initializer_params, load_variable_config, mix_seed
from ..utils import compress as _compress
from ..utils import log, metrics


class _Shard:
    def __init__(self, lib, dtype, dim, shard_id, shard_num, variable_id):
        self.lib, self.dim, self.dtype = lib, dim, dtype
        self.h = lib.exb_var_create(DTYPES[dtype], dim, 0, shard_id, shard_num, 1)
        self.np_dt = np.float32 if dtype == "float32" else np.float64
        self.shard_id, self.shard_num, self.variable_id = shard_id, shard_num, variable_id
        self.init = None      # (kind, p0, p1, p2, mixed_seed): value served for rows that were never trained

    def set_init(self, init):
        self.init = tuple(init)
        self.lib.exb_var_set_initializer(self.h, int(init[0]), float(init[1]), float(init[2]), float(init[3]), int(init[4]))

    def pull(self, local_ids):
        ids = np.ascontiguousarray(local_ids, dtype=np.uint64)
        out = np.empty((ids.size, self.dim), dtype=self.np_dt)
        if ids.size:
            self.lib.exb_var_pull(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.lib.exb_var_destroy(self.h)
            self.h = None


class ServingNode:
    def __init__(self, master_endpoint="", bind_ip="127.0.0.1", config="", port=0):
        self.lib = _native.core()
        self.models = {}        # sign -> {"status","error","uri","shard_num","shards":{(vid,shard):_Shard},"variables":[...]}
        self.lock = threading.RLock()
        self._load_lock = threading.Lock()
        self.bind_ip = bind_ip or "127.0.0.1"
        node = self

        class H(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):
                pass

            def _send(self, code, body=b"", ctype="application/json"):
                if isinstance(body, (dict, list)):
                    body = json.dumps(body).encode()
                self.send_response(code)
                # server.message_compress (reference: RpcView compress, snappy/lz4/zlib): row payloads are
                # compressed with the first codec the peer accepts (utils/compress.py: native lz4, zlib)
                if ctype == "application/octet-stream" and len(body) >= 4096:
                    accept = self.headers.get("Accept-Encoding", "")
                    for enc in ("lz4", "deflate"):
                        if enc in accept:
                            body = _compress.compress(body, enc)
                            self.send_header("Content-Encoding", enc)
                            break
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def _body(self):
                n = int(self.headers.get("Content-Length", 0))
                data = self.rfile.read(n) if n else b""
                enc = self.headers.get("Content-Encoding", "")
                if enc in ("deflate", "lz4"):
                    data = _compress.decompress(data, enc)
                return data

            def do_GET(self):
                u = urlparse(self.path)
                if u.path == "/health":
                    return self._send(200, {"ok": True, "node_id": node.node_id})
                if u.path == "/models":
                    return self._send(200, node.describe())
                if u.path == "/dump":
                    q = parse_qs(u.query)
                    data = node.dump_shard(q["model_sign"][0], int(q["variable_id"][0]), int(q["shard_id"][0]))
                    return self._send(200 if data is not None else 404, data or b"", "application/octet-stream")
                if u.path == "/metrics":
                    return self._send(200, metrics.render().encode(), "text/plain")
                self._send(404, {"error": "not found"})

            def do_POST(self):
                u = urlparse(self.path)
                if u.path == "/pull":
                    q = parse_qs(u.query)
                    with metrics.timed_request("read_only_pull"):
                        out = node.pull(q["model_sign"][0], int(q["variable_id"][0]), int(q["shard_id"][0]),
                                        np.frombuffer(self._body(), dtype=np.uint64))
                    if out is None:
                        return self._send(404, {"error": "no such model/shard"})
                    return self._send(200, out.tobytes(), "application/octet-stream")
                if u.path == "/models":
                    req = json.loads(self._body() or b"{}")
                    node.load_model_async(req)
                    return self._send(202, {"accepted": True})
                if u.path == "/shutdown":
                    self._send(200, {"ok": True})
                    threading.Thread(target=node.shutdown, daemon=True).start()
                    return
                self._send(404, {"error": "not found"})

            def do_DELETE(self):
                u = urlparse(self.path)
                if u.path.startswith("/models/"):
                    ok = node.delete_model(u.path[len("/models/"):])
                    return self._send(200 if ok else 404, {"deleted": ok})
                self._send(404, {"error": "not found"})

        self.httpd = ThreadingHTTPServer((self.bind_ip, port), H)
        self.httpd.daemon_threads = True
        self.port = self.httpd.server_address[1]
        self.master = None
        self.node_id = 0
        if master_endpoint:
            from ..master import MasterClient
            self.master = MasterClient(master_endpoint)
            self.node_id = self.master.generate_id("node")
            self.master.tree_node_add("nodes/%d" % self.node_id,
                                      json.dumps({"endpoint": self.endpoint, "node_id": self.node_id}), ephemeral=True)
        log.set_id("SERVER", self.node_id)

    @property
    def endpoint(self):
        return "%s:%d" % (self.bind_ip, self.port)

    def serve_forever(self):
        self.httpd.serve_forever(poll_interval=0.05)

    def shutdown(self):
        if self.master is not None:
            try:
                self.master.tree_node_del("nodes/%d" % self.node_id)
                self.master.close()
            except Exception:
                pass
        self.httpd.shutdown()
        self.httpd.server_close()
        with self.lock:
            for m in self.models.values():
                for s in m["shards"].values():
                    s.close()
            self.models = {}

    # ---- model management
    def describe(self):
        with self.lock:
            return {sign: {"status": m["status"], "error": m["error"], "uri": m["uri"], "shard_num": m["shard_num"],
                           "shards": sorted(set(k[1] for k in m["shards"]))} for sign, m in self.models.items()}

    def load_model_async(self, req):
        sign = req["model_sign"]
        with self.lock:
            m = self.models.get(sign)
            if m is None or int(m["shard_num"]) != int(req["shard_num"]):
                m = self.models[sign] = {"status": "LOADING", "error": "", "uri": req.get("model_uri", ""),
                                         "shard_num": int(req["shard_num"]), "shards": {}, "variables": [], "pending": 0}
            # further requests for the same model ADD shards (a restored node takes over several shard replicas,
            # each streamed from a different peer); the model is NORMAL again when all of them have landed
            m["status"] = "LOADING"
            m["pending"] = m.get("pending", 0) + 1
        threading.Thread(target=self._load, args=(req,), daemon=True).start()

    def _load(self, req):
        sign = req["model_sign"]
        try:
            with self._load_lock:        # loads of one node run one at a time
                if req.get("peer"):
                    self._load_from_peer(req)
                else:
                    self._load_from_fs(req)
            with self.lock:
                m = self.models[sign]
                m["pending"] = max(0, m.get("pending", 1) - 1)
                if m["pending"] == 0 and m["status"] != "ERROR":
                    m["status"] = "NORMAL"
        except Exception as e:      # surfaces through GET /models like the reference's model_error
            log.error("load %s failed: %r", sign, e)
            with self.lock:
                self.models[sign]["pending"] = max(0, self.models[sign].get("pending", 1) - 1)
                self.models[sign]["status"] = "ERROR"
                self.models[sign]["error"] = repr(e)

    def _new_shards(self, req, variables):
        sign, S = req["model_sign"], int(req["shard_num"])
        shards = {}
        for vid, v in enumerate(variables):
            for sid in req["shards"]:
                shards[(vid, int(sid))] = _Shard(self.lib, v["datatype"], int(v["embedding_dim"]), int(sid), S, vid)
        with self.lock:
            old = self.models[sign]["shards"]
            for k, sh in shards.items():
                if k in old:
                    old[k].close()
                old[k] = sh
            self.models[sign]["variables"] = variables
        return shards

    def _load_from_fs(self, req):
        uri, S = req["model_uri"], int(req["shard_num"])
        meta = read_model_meta(uri)
        variables = meta["variables"]
        shards = self._new_shards(req, variables)
        mine = set(int(s) for s in req["shards"])
        gvid = {}   # (storage_name, vid-in-storage) -> global variable index
        per_storage = {}
        for i, v in enumerate(variables):
            k = per_storage.get(v["storage_name"], 0)
            gvid[(v["storage_name"], k)] = i
            per_storage[v["storage_name"]] = k + 1
        for st in sorted(set(v["storage_name"] for v in variables)):
            sdir = os.path.join(uri, st)
            if not os.path.isdir(sdir):
                continue
            for fn in sorted(os.listdir(sdir)):
                if not fn.startswith("model_"):
                    continue
                for rec in iter_shard_file(os.path.join(sdir, fn)):
                    if rec[0] == "header":
                        hdr = rec[1]
                        vid = gvid[(st, hdr["variable_id"])]
                        cfg = load_variable_config(hdr["config"])
                        if "initializer" in cfg:    # rows never trained are served with the initializer value
                            kind, p, seed = initializer_params(cfg["initializer"])
                            for sid in mine:
                                shards[(vid, sid)].set_init((kind, p[0], p[1], p[2], mix_seed(seed, vid)))
                        continue
                    _, hdr, gid, w, s = rec
                    vid = gvid[(st, hdr["variable_id"])]
                    sh = (gid % np.uint64(S)).astype(np.int64)
                    for sid in mine:
                        m = sh == sid
                        if not m.any():
                            continue
                        local = np.ascontiguousarray(gid[m] // np.uint64(S))
                        ww = np.ascontiguousarray(w[m])
                        self.lib.exb_var_set_weights(shards[(vid, sid)].h, local.ctypes.data, local.size,
                                                     ww.ctypes.data, None, 0)

    def _load_from_peer(self, req):
        """coordinated restore: stream the shards from a live replica (GET /dump)"""
        import urllib.request
        variables = req["variables"]
        shards = self._new_shards(req, variables)
        for (vid, sid), shard in shards.items():
            url = "http://%s/dump?model_sign=%s&variable_id=%d&shard_id=%d" % (req["peer"], req["model_sign"], vid, sid)
            data = urllib.request.urlopen(url, timeout=60).read()
            n, hl = (int(x) for x in np.frombuffer(data[:16], dtype=np.uint64))
            hdr = json.loads(data[16:16 + hl].decode())
            if hdr.get("init"):
                shard.set_init(hdr["init"])
            data = data[16 + hl:]
            ids = np.frombuffer(data[:8 * n], dtype=np.uint64).copy()
            w = np.frombuffer(data[8 * n:], dtype=shard.np_dt).reshape(n, shard.dim).copy()
            if n:
                self.lib.exb_var_set_weights(shard.h, ids.ctypes.data, n, w.ctypes.data, None, 0)

    def dump_shard(self, sign, vid, sid):
        with self.lock:
            m = self.models.get(sign)
            shard = m["shards"].get((vid, sid)) if m else None
        if shard is None:
            return None
        n = int(self.lib.exb_var_num_items(shard.h))
        ids = np.empty(max(n, 1), dtype=np.uint64)
        cur = ctypes.c_uint64(0)
        k = int(self.lib.exb_var_read_indices(shard.h, ctypes.byref(cur), ids.ctypes.data, max(n, 1))) if n else 0
        ids = ids[:k]
        hdr = json.dumps({"init": list(shard.init) if shard.init else None}).encode()
        return (np.array([k, len(hdr)], dtype=np.uint64).tobytes() + hdr + ids.tobytes() + shard.pull(ids).tobytes())

    def delete_model(self, sign):
        with self.lock:
            m = self.models.pop(sign, None)
        if m is None:
            return False
        for s in m["shards"].values():
            s.close()
        return True

    def pull(self, sign, vid, sid, local_ids):
        with self.lock:
            m = self.models.get(sign)
            shard = m["shards"].get((vid, sid)) if (m and m["status"] == "NORMAL") else None
        if shard is None:
            return None
        return shard.pull(local_ids)


def main(argv=None):
    """``python -m openembedding_b200.serving.node --master_endpoint ip:port [--restore]``
    (flags mirror openembedding/entry/server.cc:7-19)"""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="")
    ap.add_argument("--config_file", default="")
    ap.add_argument("--master_endpoint", default="")
    ap.add_argument("--rpc_bind_ip", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("--restore", action="store_true")
    ap.add_argument("--enable_metrics", action="store_true")
    ap.add_argument("--metrics_ip", default="0.0.0.0")
    ap.add_argument("--metrics_port", type=int, default=8001)
    a = ap.parse_args(argv)
    cfg = open(a.config_file).read() if a.config_file else a.config
    node = ServingNode(master_endpoint=a.master_endpoint, bind_ip=a.rpc_bind_ip, config=cfg, port=a.port)
    if a.enable_metrics:
        metrics.start_exposer(a.metrics_ip, a.metrics_port)
    log.info("serving node %d listening on %s", node.node_id, node.endpoint)
    serving = threading.Thread(target=node.serve_forever, daemon=True)
    serving.start()                      # the restore below POSTs the shard loads to this very node
    if a.restore and a.master_endpoint:
        from .controller import ModelController
        restored = ModelController(a.master_endpoint).restore_node(node.node_id, node.endpoint)
        log.info("restored %d shard replica(s) of a dead node", len(restored))
    try:
        while serving.is_alive():
            serving.join(1.0)
    except KeyboardInterrupt:
        node.shutdown()


if __name__ == "__main__":
    main()
