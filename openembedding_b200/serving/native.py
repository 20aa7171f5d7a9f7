"""In-process serving of an exported model: no master, no nodes, no sockets.

Reference: pico-ps ``NativePS`` / ``NativePullHandler`` / ``NativeLoadHandler``
(pico-ps/pico-ps/native_ps/*: "single-process serving, no RPC threads") -- the predictor links
the tables into its own process. ``NativeModel(uri)`` loads every shard of a model written by
``save_server_model`` / ``model.save(...)`` into the native CPU shard engine of this process and
answers ``pull`` directly; rows that were never trained are served with their initializer value,
exactly like the serving nodes do.
"""
import numpy as np
import torch

from ..checkpoint import read_model_meta
from .node import ServingNode


class NativeModel:
    def __init__(self, model_uri, shard_num=-1):
        meta = read_model_meta(model_uri)
        self.model_sign = meta["model_sign"]
        self.variables = meta["variables"]
        if shard_num is None or shard_num <= 0:
            shard_num = 1
        self.shard_num = int(shard_num)
        self._node = ServingNode(master_endpoint="", port=0)        # never serves: used for its loader / shards
        req = {"model_sign": self.model_sign, "model_uri": model_uri, "shard_num": self.shard_num,
               "shards": list(range(self.shard_num))}
        with self._node.lock:
            self._node.models[self.model_sign] = {"status": "LOADING", "error": "", "uri": model_uri,
                                                  "shard_num": self.shard_num, "shards": {}, "variables": []}
        self._node._load_from_fs(req)
        self._node.models[self.model_sign]["status"] = "NORMAL"

    def embedding_dim(self, variable_id):
        return int(self.variables[variable_id]["embedding_dim"])

    def pull(self, variable_id, indices):
        """indices: int tensor / array of any shape -> torch tensor ``indices.shape + (dim,)``"""
        ids = torch.as_tensor(indices).reshape(-1).to(torch.int64).numpy().astype(np.uint64)
        dim = self.embedding_dim(variable_id)
        dt = np.float32 if self.variables[variable_id]["datatype"] == "float32" else np.float64
        out = np.empty((ids.size, dim), dtype=dt)
        S = np.uint64(self.shard_num)
        shard_of = (ids % S).astype(np.int64)
        for s in range(self.shard_num):
            sel = np.nonzero(shard_of == s)[0]
            if sel.size:
                out[sel] = self._node.pull(self.model_sign, variable_id, s, ids[sel] // S)
        return torch.from_numpy(out).reshape(tuple(torch.as_tensor(indices).shape) + (dim,))

    def close(self):
        self._node.delete_model(self.model_sign)
        self._node.httpd.server_close()
