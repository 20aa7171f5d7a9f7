"""Host-DRAM overflow tier: tables larger than HBM.

Reference counterpart: the PMem tables (openembedding/variable/PmemEmbeddingTable.h:107-417,
PmemEmbeddingItemPool.h:133-365, PmemEmbeddingOptimizerVariable.h:18-198, PersistManager.h):
all rows live in persistent memory, hot rows in a DRAM cache, pulls register an async
"touch" that promotes rows after replying, every item carries the batch id (``work_id``)
so that a checkpoint can flush exactly the rows older than the checkpoint batch, and
``persist_server_model`` writes header-only dump records (``num_items = 0``) that point at
the pool.

B200 mapping:   PMem pool  -> pinned host DRAM store (``libexb_core`` hash variable = weights
                              + optimizer state, authoritative for non-resident rows)
                DRAM cache -> the HBM table of the CUDA engine (bounded number of rows)
                promotion  -> ``prefetch(ids)``: gather from the store into a pinned staging
                              buffer, ``cudaMemcpyAsync`` + scatter kernel on a side stream
                eviction   -> write-back of dirty rows (device gather -> D2H -> store), epoch
                              ("clock") eviction of the whole cache when the budget is exceeded
                checkpoint -> ``persist``: flush rows dirtied before the checkpoint batch,
                              dump the store with the native shard-file writer, emit
                              header-only records {host_pool_path, checkpoint}

The tier is backend agnostic (the "cache" is any backend variable), so the protocol is unit
tested on the CPU backend and runs unchanged on the CUDA engine.
"""
import ctypes
import json
import os
import threading
import time

import numpy as np
import torch

from . import _native
from .config import DTYPES, initializer_params, mix_seed, optimizer_params, optimizer_state_dim

_tiers = {}          # variable_id -> TieredVariable


class CacheBudget:
    """Global cache budget in bytes (reference PersistManager: dynamic 2/3 + reserved 1/3,
    openembedding/client/Connection.cpp:87-93)."""

    def __init__(self, total_bytes):
        self.total = int(total_bytes)
        self.dynamic = self.total * 2 // 3
        self.reserved = self.total - self.dynamic
        self.used = 0
        self.lock = threading.Lock()

    def acquire(self, nbytes):
        with self.lock:
            if self.used + nbytes > self.dynamic:
                return False
            self.used += nbytes
            return True

    def release(self, nbytes):
        with self.lock:
            self.used = max(0, self.used - nbytes)


class TieredVariable:
    def __init__(self, ctx, meta, cache_rows, pool_path=None):
        self.ctx, self.meta, self.be = ctx, meta, ctx.backend
        self.lib = _native.core()
        self.cache_rows = int(cache_rows)
        self.pool_path = pool_path
        self.dim = meta.dim
        self.np_dt = np.float32 if meta.dtype == "float32" else np.float64
        # the store holds THIS RANK's shard of the table (hash: any id)
        self.store = self.lib.exb_var_create(DTYPES[meta.dtype], meta.dim, 0, 0, 1, 1)
        self._sync_store_config()
        self.resident = {}       # global id -> work_id of the last update (0 = clean)
        self.work_id = 1         # batch counter (reference: next_work / work_id stamps)
        self.checkpoint = 0
        self.pending_persist = False
        self.stats = {"hits": 0, "misses": 0, "evictions": 0, "flushes": 0, "writebacks": 0}
        self._lock = threading.Lock()
        _tiers[meta.variable_id] = self

    # ---- config mirrored into the store so that state layouts agree
    def _sync_store_config(self):
        kind, p, seed = initializer_params(self.meta.initializer)
        self.lib.exb_var_set_initializer(self.store, kind, p[0], p[1], p[2], mix_seed(seed, self.meta.variable_id))
        ok, op = optimizer_params(self.meta.optimizer)
        self.lib.exb_var_set_optimizer(self.store, ok, (ctypes.c_double * 8)(*op), 8)
        self.sd = optimizer_state_dim(self.meta.optimizer, self.dim)

    def _owned(self, ids):
        owner = (self.meta.shard_base + ids % self.meta.shard_num) % self.ctx.world
        return ids[owner == self.ctx.rank]

    # ---- promotion
    def prefetch(self, ids):
        """Make the rows of `ids` (any shape, int64) that this rank owns resident in the cache."""
        self._sync_store_config()
        ids = torch.unique(ids.reshape(-1).to("cpu", torch.int64))
        if self.ctx.world > 1:
            ids = self._allgather_ids(ids)
        mine = self._owned(ids).numpy().astype(np.uint64)
        with self._lock:
            have = np.fromiter(self.resident.keys(), dtype=np.uint64, count=len(self.resident))
            miss = mine[~np.isin(mine, have)] if have.size else mine
            self.stats["hits"] += int(mine.size - miss.size)
            self.stats["misses"] += int(miss.size)
            if miss.size == 0:
                return 0
            if len(self.resident) + miss.size > self.cache_rows:
                self._evict_all_locked()
                miss = mine            # everything needed now has to be (re)loaded
            w = np.empty((miss.size, self.dim), dtype=self.np_dt)
            s = np.empty((miss.size, max(self.sd, 1)), dtype=self.np_dt)
            self.lib.exb_var_get_weights(self.store, miss.ctypes.data, miss.size, w.ctypes.data,
                                         s.ctypes.data if self.sd else None)
            self.be.load_rows(self.meta, miss, w, s[:, :self.sd] if self.sd else np.empty((miss.size, 0), self.np_dt))
            for i in miss.tolist():
                self.resident[i] = 0
        return int(miss.size)

    def _allgather_ids(self, ids):
        import torch.distributed as dist
        objs = [None] * self.ctx.world
        dist.all_gather_object(objs, ids, group=self.ctx.group)
        return torch.unique(torch.cat(objs))

    # ---- write-back / eviction
    def _cached_rows(self, gids):
        """(ids, weights, states) of resident rows read back from the cache backend (direct gather by id)"""
        ids = np.ascontiguousarray(np.fromiter((int(x) for x in gids), dtype=np.uint64, count=len(gids)))
        if ids.size == 0:
            return ids, np.empty((0, self.dim), self.np_dt), np.empty((0, self.sd), self.np_dt)
        w, s = self.be.read_rows(self.meta, ids)
        return ids, np.asarray(w), np.asarray(s)

    def _writeback_locked(self, gids):
        if len(gids) == 0:
            return
        self.be.synchronize()
        ids, w, s = self._cached_rows(gids)
        if ids.size:
            ids = np.ascontiguousarray(ids); w = np.ascontiguousarray(w, dtype=self.np_dt)
            s = np.ascontiguousarray(s, dtype=self.np_dt)
            self.lib.exb_var_set_weights(self.store, ids.ctypes.data, ids.size, w.ctypes.data,
                                         s.ctypes.data if self.sd else None, self.sd * w.itemsize)
            self.stats["writebacks"] += int(ids.size)

    def _evict_all_locked(self):
        dirty = [i for i, wid in self.resident.items() if wid > 0]
        self._writeback_locked(dirty)
        self.stats["evictions"] += len(self.resident)
        self.be.clear(self.meta)
        self.resident = {}
        self.stats["flushes"] += 1

    def flush(self, before_work_id=None):
        """write back rows dirtied before `before_work_id` (all dirty rows if None)"""
        with self._lock:
            dirty = [i for i, wid in self.resident.items()
                     if wid > 0 and (before_work_id is None or wid <= before_work_id)]
            self._writeback_locked(dirty)
            for i in dirty:
                self.resident[i] = 0

    # ---- bookkeeping called by the variable wrapper
    def mark_updated(self, ids):
        """collective when world > 1: a row this rank owns is dirtied by ANY rank's push"""
        ids = torch.unique(ids.reshape(-1).to("cpu", torch.int64))
        if self.ctx.world > 1:
            ids = self._allgather_ids(ids)
        ids = self._owned(ids).tolist()
        with self._lock:
            for i in ids:
                if i in self.resident:
                    self.resident[i] = self.work_id

    def next_work(self):
        self.work_id += 1

    # ---- lightweight checkpoint (reference: start_commit_checkpoint / flush_committing_checkpoint)
    def persist(self, pool_dir, window=0):
        """Flush what the checkpoint needs and dump the store. `window` batches may stay
        un-flushed (reference persist_pending_window)."""
        ckpt = max(0, self.work_id - int(window))
        self.flush(before_work_id=ckpt)
        os.makedirs(pool_dir, exist_ok=True)
        fn = os.path.join(pool_dir, "pool_v%d_r%d" % (self.meta.variable_id, self.ctx.rank))
        self._dump_store(fn)
        self.checkpoint = ckpt
        self.pool_path = pool_dir
        return {"host_pool_path": pool_dir, "checkpoint": int(ckpt)}

    def _dump_store(self, fn):
        lib = self.lib
        w = lib.exb_fw_open(fn.encode())
        n = int(lib.exb_var_num_items(self.store))
        itemsize = np.dtype(self.np_dt).itemsize
        lib.exb_fw_header(w, self.meta.variable_id, DTYPES[self.meta.dtype], self.dim, self.meta.vocab, b"", 0,
                          0, 1, self.sd * itemsize, n)
        cursor = ctypes.c_uint64(0)
        blk = 1 << 15
        while True:
            idx = np.empty(blk, dtype=np.uint64)
            k = int(lib.exb_var_read_indices(self.store, ctypes.byref(cursor), idx.ctypes.data, blk))
            if k == 0:
                break
            idx = idx[:k]
            ww = np.empty((k, self.dim), dtype=self.np_dt)
            ss = np.empty((k, max(self.sd, 1)), dtype=self.np_dt)
            lib.exb_var_get_weights(self.store, idx.ctypes.data, k, ww.ctypes.data, ss.ctypes.data if self.sd else None)
            ss = np.ascontiguousarray(ss[:, :self.sd])
            lib.exb_fw_block(w, k, idx.ctypes.data, ww.ctypes.data, ww.nbytes, ss.ctypes.data if ss.nbytes else None, ss.nbytes)
        lib.exb_fw_close(w)

    def restore(self, pool_dir):
        from .checkpoint import iter_shard_file
        fn = os.path.join(pool_dir, "pool_v%d_r%d" % (self.meta.variable_id, self.ctx.rank))
        self.lib.exb_var_clear(self.store)
        self._sync_store_config()
        with self._lock:
            self.be.clear(self.meta)
            self.resident = {}
        if not os.path.exists(fn):
            return
        for rec in iter_shard_file(fn):
            if rec[0] != "block":
                continue
            _, hdr, gid, w, s = rec
            gid = np.ascontiguousarray(gid); w = np.ascontiguousarray(w); s = np.ascontiguousarray(s)
            self.lib.exb_var_set_weights(self.store, gid.ctypes.data, gid.size, w.ctypes.data,
                                         s.ctypes.data if s.size else None, s.shape[1] * w.itemsize if s.size else 0)
        self.pool_path = pool_dir

    def close(self):
        if self.store:
            self.lib.exb_var_destroy(self.store)
            self.store = None
        _tiers.pop(self.meta.variable_id, None)


# ------------------------------------------------------------------------------------------
def make_tiered(ctx, meta, cache_rows=None):
    """Attach a host tier to a (hash) variable. cache_rows defaults to the EnvConfig
    ``server.cache_size`` (MB) budget divided by the row footprint."""
    if cache_rows is None:
        mb = int(ctx.env["server"]["cache_size"])
        row_bytes = (meta.dim + optimizer_state_dim(meta.optimizer, meta.dim)) * 4 + 16
        cache_rows = max(1024, mb * (1 << 20) * 2 // 3 // row_bytes)
    root = ctx.env["server"].get("host_tier_root_path") or ctx.env["server"].get("pmem_pool_root_path") or None
    return TieredVariable(ctx, meta, cache_rows, pool_path=root)


def tier_of(meta):
    return _tiers.get(meta.variable_id)


def should_persist(ctx):
    """True when some tier has accumulated enough dirty rows that a lightweight checkpoint is
    worthwhile (reference piggy-backs this on pull responses, EmbeddingPullOperator.cpp:184-189)."""
    for t in _tiers.values():
        dirty = sum(1 for wid in t.resident.values() if wid > t.checkpoint)
        if dirty * 2 > t.cache_rows:
            return True
    return False


def persist_model(ctx, filepath, persist_pending_window=0):
    """Collective lightweight checkpoint: header-only dump + per-tier pool files."""
    from . import checkpoint as ck
    pool = filepath.rstrip("/") + ".pool"
    extra = {}
    for t in _tiers.values():
        extra = t.persist(pool, persist_pending_window)
    if filepath.startswith("mem://null/"):
        return
    ck.save_model(ctx, filepath, include_optimizer=True, persist=extra or {"host_pool_path": pool, "checkpoint": 0})
    if ctx.rank == 0:
        with open(os.path.join(filepath, "persist.json"), "w") as fh:
            json.dump({"pool": pool, "time": time.time(),
                       "tiers": {str(k): {"checkpoint": t.checkpoint} for k, t in _tiers.items()}}, fh)
    ctx.barrier()


def restore_model(ctx, filepath):
    """Collective: re-open the pools named by a ``persist_model`` checkpoint."""
    from . import checkpoint as ck
    meta = ck.read_model_meta(filepath)
    mine = ck.model_meta_dict(ctx)["variables"]
    if meta["variables"] != mine:
        raise ValueError("model meta not match")
    pool = filepath.rstrip("/") + ".pool"
    pj = os.path.join(filepath, "persist.json")
    if os.path.exists(pj):
        pool = json.load(open(pj))["pool"]
    for t in _tiers.values():
        t.restore(pool)
    ctx.barrier()
