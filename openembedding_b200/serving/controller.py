"""Model controller + REST front end.

Reference: ``ModelController`` (openembedding/client/ModelController.cpp:47-162: create /
delete model with a master-tree lock, CREATING -> NORMAL status, node listing / shutdown)
and the brpc HTTP ``controller`` daemon (openembedding/entry/controller.cc:54-259):
``POST/GET/DELETE /models[/sign]``, ``GET/DELETE /nodes[/id]``, default port 8010.
Replica placement follows openembedding/client/Model.cpp:153-186 (shard s, replica r ->
node (s + r) mod N).
"""
import json
import threading
import time
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from urllib.parse import urlparse

from ..checkpoint import read_model_meta
from ..master import MasterClient

MODEL_STATUS = ("CREATING", "NORMAL", "DELETING", "ERROR", "LOADING")   # pico-ps model/Model.h ModelStatus


def _http(method, url, body=None, timeout=30):
    req = urllib.request.Request(url, data=json.dumps(body).encode() if body is not None else None, method=method,
                                 headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(req, timeout=timeout) as r:
        data = r.read()
    try:
        return json.loads(data)
    except Exception:
        return data


class ModelController:
    def __init__(self, master_endpoint):
        self.master = MasterClient(master_endpoint)

    # ---- nodes
    def nodes(self):
        out = {}
        for name in self.master.tree_node_sub("nodes"):
            v = self.master.tree_node_get("nodes/" + name)
            if v:
                out[int(name)] = json.loads(v)
        return out

    def show_node(self, node_id):
        n = self.nodes().get(int(node_id))
        if n is None:
            return None
        try:
            n["models"] = _http("GET", "http://%s/models" % n["endpoint"])
            n["status"] = "RUNNING"
        except Exception as e:
            n["status"], n["error"] = "DEAD", repr(e)
        return n

    def shutdown_node(self, node_id):
        n = self.nodes().get(int(node_id))
        if n is None:
            return False
        try:
            _http("POST", "http://%s/shutdown" % n["endpoint"], {})
        except Exception:
            pass
        self.master.tree_node_del("nodes/%d" % int(node_id))
        return True

    # ---- models
    def models(self):
        out = {}
        for name in self.master.tree_node_sub("models"):
            v = self.master.tree_node_get("models/" + name)
            if v:
                out[name] = json.loads(v)
        return out

    def show_model(self, sign):
        v = self.master.tree_node_get("models/" + sign)
        return json.loads(v) if v else None

    def create_model(self, model_uri, replica_num=1, shard_num=-1, wait=True, timeout=600):
        meta = read_model_meta(model_uri)
        sign = meta["model_sign"]
        self.master.acquire_lock("model/" + sign)
        try:
            if self.show_model(sign) is not None:
                raise ValueError("model already exists: " + sign)
            nodes = self.nodes()
            if not nodes:
                raise RuntimeError("no serving node registered")
            ids = sorted(nodes)
            S = len(ids) if shard_num in (-1, 0, None) else int(shard_num)
            R = max(1, min(int(replica_num), len(ids)))
            # rotating cursor over the nodes (Model.cpp:153-186: start = ino.fetch_add(shard_num * replica_num),
            # node = (shard * replica_num + i + start) % n): consecutive models -- and single-shard models in
            # particular -- do not pile up on the first node. The cursor lives in the master, so every controller
            # of the cluster advances the same one (the reference's is per process).
            start = self.master.advance_counter("placement_cursor", S * R)
            placement = {}
            for s_ in range(S):
                reps = []
                for r in range(R):
                    nid = ids[(s_ * R + r + start) % len(ids)]
                    if nid in reps:                       # R does not divide the ring evenly: next free node
                        nid = next(x for x in ids[(ids.index(nid) + 1):] + ids if x not in reps)
                    reps.append(nid)
                placement[str(s_)] = reps
            rec = {"model_sign": sign, "model_uri": model_uri, "model_status": "CREATING", "model_error": "",
                   "variables": meta["variables"], "shard_num": S, "replica_num": R, "placement": placement}
            self.master.tree_node_set("models/" + sign, json.dumps(rec))
        finally:
            self.master.release_lock("model/" + sign)
        per_node = {}
        for s, reps in placement.items():
            for nid in reps:
                per_node.setdefault(nid, []).append(int(s))
        for nid, shards in per_node.items():
            _http("POST", "http://%s/models" % nodes[nid]["endpoint"],
                  {"model_sign": sign, "model_uri": model_uri, "shards": shards, "shard_num": S})
        if wait:
            self._wait_normal(sign, per_node, nodes, timeout)
        return sign

    def _wait_normal(self, sign, per_node, nodes, timeout):
        t0 = time.time()
        status, err = "CREATING", ""
        while time.time() - t0 < timeout:
            states = []
            for nid in per_node:
                try:
                    m = _http("GET", "http://%s/models" % nodes[nid]["endpoint"]).get(sign, {})
                    states.append((m.get("status", "LOADING"), m.get("error", "")))
                except Exception as e:
                    states.append(("ERROR", repr(e)))
            if any(s == "ERROR" for s, _ in states):
                status, err = "ERROR", "; ".join(e for s, e in states if s == "ERROR")
                break
            if all(s == "NORMAL" for s, _ in states):
                status = "NORMAL"
                break
            time.sleep(0.05)
        rec = self.show_model(sign)
        rec["model_status"], rec["model_error"] = status, err
        self.master.tree_node_set("models/" + sign, json.dumps(rec))
        if status != "NORMAL":
            raise RuntimeError("create_model %s: %s %s" % (sign, status, err))

    def delete_model(self, sign):
        rec = self.show_model(sign)
        if rec is None:
            return False
        nodes = self.nodes()
        for reps in rec["placement"].values():
            for nid in reps:
                if nid in nodes:
                    try:
                        _http("DELETE", "http://%s/models/%s" % (nodes[nid]["endpoint"], sign))
                    except Exception:
                        pass
        self.master.tree_node_del("models/" + sign)
        return True

    # ---- HA: a fresh node takes over the shards of a dead one (server --restore,
    #      pico-ps service/Service.cpp:237-313 restore_storages)
    def restore_node(self, new_id, new_endpoint):
        """A fresh node takes the place of ONE dead node: every shard replica the dead node held (in every
        model) is re-created on the new node, streamed from a live replica when there is one, else re-read
        from the model uri (reference: CoordinatedRestoreController, one replacement per dead node)."""
        live = self.nodes()
        models = self.models()
        dead = None
        for rec in models.values():
            for reps in rec["placement"].values():
                for nid in reps:
                    if nid not in live and nid != new_id:
                        dead = nid if dead is None else min(dead, nid)
        if dead is None:
            return []
        restored = []
        for sign, rec in models.items():
            changed = False
            for s, reps in rec["placement"].items():
                for i, nid in enumerate(reps):
                    if nid != dead:
                        continue
                    peers = [p for p in reps if p in live and p != new_id]
                    req = {"model_sign": sign, "model_uri": rec["model_uri"], "shards": [int(s)],
                           "shard_num": rec["shard_num"], "variables": rec["variables"]}
                    if peers:
                        req["peer"] = live[peers[0]]["endpoint"]     # stream from a live replica
                    _http("POST", "http://%s/models" % new_endpoint, req)
                    reps[i] = new_id
                    changed = True
                    restored.append((sign, int(s)))
            if changed:
                self.master.tree_node_set("models/" + sign, json.dumps(rec))
        return restored


def serve(master_endpoint, port=8010, bind_ip="0.0.0.0"):
    ctl = ModelController(master_endpoint)

    class H(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def log_message(self, *a):
            pass

        def _send(self, code, obj):
            body = json.dumps(obj).encode()
            self.send_response(code)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            p = urlparse(self.path).path.rstrip("/")
            if p == "/models":
                return self._send(200, ctl.models())
            if p.startswith("/models/"):
                m = ctl.show_model(p[len("/models/"):])
                return self._send(200 if m else 404, m or {"error": "no such model"})
            if p == "/nodes":
                return self._send(200, {str(k): v for k, v in ctl.nodes().items()})
            if p.startswith("/nodes/"):
                n = ctl.show_node(p[len("/nodes/"):])
                return self._send(200 if n else 404, n or {"error": "no such node"})
            self._send(404, {"error": "not found"})

        def do_POST(self):
            p = urlparse(self.path).path.rstrip("/")
            n = int(self.headers.get("Content-Length", 0))
            body = json.loads(self.rfile.read(n) or b"{}") if n else {}
            if p == "/models":
                try:
                    sign = ctl.create_model(body["model_uri"], body.get("replica_num", 3), body.get("shard_num", -1),
                                            wait=bool(body.get("wait", True)))
                    return self._send(200, {"model_sign": sign})
                except Exception as e:
                    return self._send(500, {"error": repr(e)})
            self._send(404, {"error": "not found"})

        def do_DELETE(self):
            p = urlparse(self.path).path.rstrip("/")
            if p.startswith("/models/"):
                ok = ctl.delete_model(p[len("/models/"):])
                return self._send(200 if ok else 404, {"deleted": ok})
            if p.startswith("/nodes/"):
                ok = ctl.shutdown_node(p[len("/nodes/"):])
                return self._send(200 if ok else 404, {"shutdown": ok})
            self._send(404, {"error": "not found"})

    httpd = ThreadingHTTPServer((bind_ip, port), H)
    httpd.daemon_threads = True
    return httpd, ctl


def main(argv=None):
    """``python -m openembedding_b200.serving.controller --master_endpoint ip:port --port 8010``"""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--master_endpoint", required=True)
    ap.add_argument("--port", type=int, default=8010)
    ap.add_argument("--bind_ip", default="0.0.0.0")
    a = ap.parse_args(argv)
    httpd, _ = serve(a.master_endpoint, a.port, a.bind_ip)
    try:
        httpd.serve_forever()
    except KeyboardInterrupt:
        httpd.shutdown()


if __name__ == "__main__":
    main()
