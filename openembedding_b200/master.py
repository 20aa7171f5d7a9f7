"""Control plane: ``Master`` (rendezvous KV tree) and ``Server`` (standalone shard host).

Reference: the TCP master is a single-thread poll() KV tree with ephemeral nodes and
watchers (pico-core rpc/Master.cpp:139-304) used for rank generation, barriers, locks
and the model registry; ``Server`` joins a job as a dedicated PS process
(openembedding/entry/server.cc:25-60, py_api.cc:164-215).

B200 design: inside one NVSwitch box every rank hosts its own shards in HBM, so the
data plane needs no server process. The control plane keeps the same verbs on top of a
``torch.distributed.TCPStore`` (tree paths are store keys): ``Master`` owns the store
server and hands out ``endpoint``; ``MasterClient`` provides tree_node_*/barrier/
acquire_lock/generate_id; ``Server`` is the out-of-job serving host (see
``serving/``).
"""
import datetime
import socket
import threading
import time
import uuid


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class Master:
    """In-process master; ``endpoint`` is what workers put in ``flags.master_endpoint``."""

    def __init__(self, bind_ip="127.0.0.1", port=0):
        from torch.distributed import TCPStore
        self.ip = bind_ip or "127.0.0.1"
        self.port = port or _free_port()
        self._store = TCPStore(self.ip, self.port, is_master=True, wait_for_workers=False,
                               timeout=datetime.timedelta(seconds=3600))
        self._closed = False

    @property
    def endpoint(self):
        return "%s:%d" % (self.ip, self.port)

    def client(self):
        return MasterClient(self.endpoint, store=self._store)

    def finalize(self):
        self._closed = True
        self._store = None

    join = finalize


class MasterClient:
    """Tree/barrier/lock/id verbs of pico-core MasterClient (rpc/MasterClient.h:56-160)."""

    def __init__(self, endpoint, root_path="/openembedding", store=None, timeout=3600):
        from torch.distributed import TCPStore
        ip, port = endpoint.rsplit(":", 1)
        self.root = root_path.rstrip("/")
        self._store = store or TCPStore(ip, int(port), is_master=False,
                                        timeout=datetime.timedelta(seconds=timeout))
        self._session = uuid.uuid4().hex

    def _k(self, path):
        return self.root + "/" + path.strip("/")

    # -- tree
    def tree_node_add(self, path, value="", ephemeral=False):
        key = self._k(path)
        # compare_set with empty expected value creates the key only if it did not exist
        got = self._store.compare_set(key, "", "v:" + value)
        ok = got.decode() == "v:" + value
        if ok:
            self._store.add(self._k("__children__/" + path.strip("/").rsplit("/", 1)[0] if "/" in path.strip("/") else "__children__/"), 1)
            idx_key = self._k("__index__")
            self._store.append(idx_key, key + "\n") if hasattr(self._store, "append") else None
        return ok

    def tree_node_set(self, path, value):
        self._store.set(self._k(path), "v:" + value)
        if hasattr(self._store, "append"):
            self._store.append(self._k("__index__"), self._k(path) + "\n")
        return True

    def tree_node_get(self, path, default=None):
        key = self._k(path)
        if not self._store.check([key]):
            return default
        v = self._store.get(key).decode()
        if v == "v:\x00deleted":
            return default
        return v[2:] if v.startswith("v:") else v

    def tree_node_del(self, path):
        key = self._k(path)
        if not self._store.check([key]):
            return False
        self._store.set(key, "v:\x00deleted")
        return True

    def tree_node_sub(self, path):
        prefix = self._k(path).rstrip("/") + "/"
        idx = self._k("__index__")
        if not self._store.check([idx]):
            return []
        keys = sorted(set(k for k in self._store.get(idx).decode().split("\n") if k.startswith(prefix)))
        out = []
        for k in keys:
            rest = k[len(prefix):]
            if "/" in rest:
                continue
            v = self._store.get(k).decode()
            if v != "v:\x00deleted":
                out.append(rest)
        return out

    # -- ids / barriers / locks
    def generate_id(self, name):
        return int(self._store.add(self._k("__id__/" + name), 1)) - 1

    def barrier(self, name, n, timeout=3600):
        key = self._k("__barrier__/" + name)
        arrived = int(self._store.add(key, 1))
        generation = (arrived - 1) // n
        target = (generation + 1) * n
        t0 = time.time()
        while int(self._store.add(key, 0)) < target:
            if time.time() - t0 > timeout:
                raise TimeoutError("master barrier " + name)
            time.sleep(0.002)

    def acquire_lock(self, name, timeout=3600):
        key = self._k("__lock__/" + name)
        t0 = time.time()
        while True:
            got = self._store.compare_set(key, "", self._session).decode()
            if got == self._session:
                return
            if got == "free":
                got = self._store.compare_set(key, "free", self._session).decode()
                if got == self._session:
                    return
            if time.time() - t0 > timeout:
                raise TimeoutError("master lock " + name)
            time.sleep(0.002)

    def release_lock(self, name):
        self._store.set(self._k("__lock__/" + name), "free")


class Server:
    """Standalone shard host for serving (see ``serving.ServingNode``).

    Training on one box never needs it: ``flags.wait_num_servers == -1`` (each worker
    hosts its shards in its own HBM) is the only training topology.
    """

    def __init__(self, master_endpoint="", bind_ip="127.0.0.1", config=""):
        from .serving.node import ServingNode
        self._node = ServingNode(master_endpoint=master_endpoint, bind_ip=bind_ip, config=config)
        self._thread = threading.Thread(target=self._node.serve_forever, daemon=True)
        self._thread.start()

    @property
    def endpoint(self):
        return self._node.endpoint

    def exit(self):
        self._node.shutdown()

    def join(self):
        self._thread.join()
