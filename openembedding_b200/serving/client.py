"""Serving-side model access: find a model variable by ``model_sign`` and pull rows
read-only from the replicas, with the reference's retry-on-failure semantics.

Reference: ``ModelManager::find_model_variable`` (openembedding/client/ModelController.cpp:
24-44), the serving branch of ``PullWeightsOp`` (openembedding/tensorflow/exb_ops.cpp:261-276:
``model_sign = uuid-floor(version)``), ``pick_one_replica`` RANDOM/ROUND_ROBIN
(pico-ps service/TableDescriptor.h:179-228) and the client retry state machine on
Timeout/NoReplica (pico-ps handler/Handler.cpp:47-106).
"""
import itertools
import json
import time
import urllib.request
import zlib

import numpy as np
import torch

from ..master import MasterClient
from ..utils import metrics


class NoReplica(RuntimeError):
    pass


class ModelVariable:
    def __init__(self, client, sign, variable_id, rec):
        self.client, self.sign, self.vid = client, sign, variable_id
        v = rec["variables"][variable_id]
        self.dim = int(v["embedding_dim"])
        self.np_dt = np.float32 if v["datatype"] == "float32" else np.float64
        self.shard_num = int(rec["shard_num"])

    def pull(self, ids, timeout=10.0):
        ids_t = torch.as_tensor(ids, dtype=torch.int64).reshape(-1)
        idn = ids_t.numpy().astype(np.uint64)
        out = np.empty((idn.size, self.dim), dtype=self.np_dt)
        shard = (idn % np.uint64(self.shard_num)).astype(np.int64)
        t0 = time.perf_counter()
        for s in np.unique(shard):
            m = shard == s
            local = np.ascontiguousarray(idn[m] // np.uint64(self.shard_num))
            out[m] = self.client._pull_shard(self.sign, self.vid, int(s), local, self.dim, self.np_dt, timeout)
        metrics.observe_wait(self.sign, "read_only_pull", (time.perf_counter() - t0) * 1e3)
        return torch.from_numpy(out).reshape(tuple(torch.as_tensor(ids).shape) + (self.dim,))


class ServingClient:
    def __init__(self, master_endpoint, policy="round_robin", message_compress=""):
        """message_compress: "", "zlib", "lz4" or "snappy" (EnvConfig server.message_compress; utils/compress.py: native
        lz4 block codec, zlib; snappy is served by the lz4 codec)."""
        self.compress = bool(message_compress)
        from ..utils import compress as _c
        self._codec = _c
        self._encoding = _c.encoding_of(message_compress) or "deflate"
        self.master = MasterClient(master_endpoint)
        self._rr = itertools.count()
        self.policy = policy
        self._dead = {}          # node_id -> time marked dead
        self._models = {}
        self._stale, self._watches = set(), {}       # model signs whose placement record changed (master watcher)

    def _model(self, sign, refresh=False):
        if refresh or sign not in self._models or sign in self._stale:
            v = self.master.tree_node_get("models/" + sign)
            if not v:
                raise KeyError("no such model: " + sign)
            self._stale.discard(sign)
            self._models[sign] = json.loads(v)
            if sign not in self._watches:     # placement changes (restore after a node loss, delete) invalidate the cache
                self._watches[sign] = self.master.watch("models/" + sign, lambda p, ver, s=sign: self._stale.add(s))
        return self._models[sign]

    def _nodes(self):
        out = {}
        for name in self.master.tree_node_sub("nodes"):
            v = self.master.tree_node_get("nodes/" + name)
            if v:
                out[int(name)] = json.loads(v)["endpoint"]
        return out

    def find_model_variable(self, model_sign, variable_id):
        rec = self._model(model_sign)
        if rec.get("model_status") != "NORMAL":
            rec = self._model(model_sign, refresh=True)
            if rec.get("model_status") != "NORMAL":
                raise RuntimeError("model %s is %s" % (model_sign, rec.get("model_status")))
        return ModelVariable(self, model_sign, variable_id, rec)

    def _pick(self, reps, key=0):
        live = [r for r in reps if time.time() - self._dead.get(r, 0) > 5.0]
        if not live:
            return None
        if self.policy == "random":
            return live[np.random.randint(len(live))]
        if self.policy == "hash":
            # replica affinity: the same (model, shard, first id) keeps hitting the same replica -- its row cache stays
            # warm -- and adding a replica moves only 1/n of the keys (jump consistent hash, utils/hashing.py)
            from ..utils.hashing import jump_consistent_hash, murmur3_fmix64
            return live[jump_consistent_hash(murmur3_fmix64(int(key)), len(live))]
        return live[next(self._rr) % len(live)]

    def _pull_shard(self, sign, vid, shard, local_ids, dim, np_dt, timeout):
        t0 = time.time()
        while True:
            rec = self._model(sign)
            reps = rec["placement"][str(shard)]
            nid = self._pick(reps, key=(int(local_ids[0]) if local_ids.size else 0) * 1000003 + shard)
            nodes = self._nodes()
            if nid is None or nid not in nodes:
                if time.time() - t0 > timeout:
                    raise NoReplica("no live replica of %s shard %d" % (sign, shard))
                if nid is not None:
                    self._dead[nid] = time.time()
                time.sleep(0.05)
                self._model(sign, refresh=True)      # placement may have been repaired by a restored node
                continue
            url = "http://%s/pull?model_sign=%s&variable_id=%d&shard_id=%d" % (nodes[nid], sign, vid, shard)
            try:
                body, hdr = local_ids.tobytes(), {}
                if self.compress:
                    hdr["Accept-Encoding"] = self._encoding
                    if len(body) >= 4096:
                        body, hdr["Content-Encoding"] = self._codec.compress(body, self._encoding), self._encoding
                req = urllib.request.Request(url, data=body, method="POST", headers=hdr)
                with urllib.request.urlopen(req, timeout=max(0.5, timeout / 4)) as r:
                    data = r.read()
                    enc = r.headers.get("Content-Encoding", "")
                    if enc in ("deflate", "lz4"):
                        data = self._codec.decompress(data, enc)
                return np.frombuffer(data, dtype=np_dt).reshape(local_ids.size, dim)
            except Exception:
                self._dead[nid] = time.time()         # handle_timeout: mark the node dead, retry elsewhere
                if time.time() - t0 > timeout:
                    raise NoReplica("pull timed out on every replica of %s shard %d" % (sign, shard))
