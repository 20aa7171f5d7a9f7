"""Host-DRAM overflow tier: tables larger than HBM.

Reference counterpart: the PMem tables (openembedding/variable/PmemEmbeddingTable.h:107-417,
PmemEmbeddingItemPool.h:133-365, PmemEmbeddingOptimizerVariable.h:18-198, PersistManager.h):
all rows live in persistent memory, hot rows in a DRAM cache, pulls register an async
"touch" that promotes rows after replying, every item carries the batch id (``work_id``)
so that a checkpoint can flush exactly the rows older than the checkpoint batch, and
``persist_server_model`` writes header-only dump records (``num_items = 0``) that point at
the pool.

Two implementations behind ``make_tiered``:

* ``GpuTieredVariable`` (CUDA engine, the product): ``csrc/cuda/host_tier.cuh``. The rows live in a PINNED
  host slab that the GPU reads / writes directly over PCIe, the id -> host-row index lives in HBM, the HBM hash
  shard of the table is cache AND residency map; promotion is ONE kernel per batch (``tier_admit_kernel``: probe,
  claim a slot, copy the host row or first-touch initialise) that ``prefetch`` enqueues -- on a side stream, one
  batch ahead, when the caller announces the next batch; eviction keeps the most recently used rows by
  ``work_id`` stamp (age histogram -> cutoff -> write back dirty victims -> in-place rebuild); the cache size
  comes out of the ``server.cache_size`` budget (``CacheBudget``, the reference's PersistManager).
* ``TieredVariable`` (CPU backend): the same protocol in python over ``libexb_core`` -- the configuration the
  no-GPU tests and the gloo plumbing run.
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


_budget = None


def cache_budget(ctx):
    """process-wide HBM-cache budget from ``server.cache_size`` (MB) -- PersistManager (Connection.cpp:87-93)"""
    global _budget
    if _budget is None:
        _budget = CacheBudget(int(ctx.env["server"]["cache_size"]) << 20)
    return _budget


class GpuTieredVariable:
    """Host-DRAM tier of one hash variable on the CUDA engine (``csrc/cuda/host_tier.cuh``)."""

    max_load, evict_to = 0.7, 0.4

    def __init__(self, ctx, meta, cache_rows, host_rows=None, pool_path=None):
        from .status import Status, StatusError
        self.ctx, self.meta, self.be = ctx, meta, ctx.backend
        self.eng = ctx.backend.engine
        self.lib = self.eng.lib
        self.pool_path = pool_path
        self.be.ensure_allocated([meta])
        info = self.eng.table_info(meta.handle)
        self.capacity = int(info["rows"])                       # slots of the HBM cache (pow2)
        self.wstride, self.sstride = int(info["wstride"]), int(info["sstride"])
        self.HR = self.wstride + self.sstride
        self.cache_rows = int(cache_rows)
        self.host_rows = int(host_rows or max(16 * self.cache_rows, 1 << 16))
        self.cache_bytes = self.capacity * self.HR * 4
        if not cache_budget(ctx).acquire(self.cache_bytes):
            raise StatusError(Status.OOM, "host tier: HBM cache of %d MB exceeds what is left of server.cache_size "
                              "(%d MB, %d MB in use)" % (self.cache_bytes >> 20, cache_budget(ctx).dynamic >> 20,
                                                         cache_budget(ctx).used >> 20))
        self.h = self.lib.exb_tier_create(self.eng.h, meta.handle, self.host_rows)
        if not self.h:
            cache_budget(ctx).release(self.cache_bytes)
            raise RuntimeError("exb_tier_create: " + self.lib.exb_cuda_last_error().decode())
        self.sd = optimizer_state_dim(meta.optimizer, meta.dim)
        self.work_id = 1
        self.clock = 1                # one tick per admitted batch: the stamps of the CLOCK eviction
        self.checkpoint = 0
        self._size_ub = 0
        self.side = torch.cuda.Stream(device=ctx.device)    # promotion stream (VariableAsyncTask analogue)
        self._ev = torch.cuda.Event()
        self._prefetched = None       # key of the batch admitted ahead of time
        meta.tiered = True
        _tiers[meta.variable_id] = self

    def on_optimizer_change(self):
        """the table's optimizer was (re)configured: state width and host-row layout follow"""
        _native.cuda_check(self.lib.exb_tier_relayout(self.h), "tier_relayout")
        info = self.eng.table_info(self.meta.handle)
        self.wstride, self.sstride = int(info["wstride"]), int(info["sstride"])
        self.HR = self.wstride + self.sstride
        self.sd = optimizer_state_dim(self.meta.optimizer, self.meta.dim)

    # ---- promotion
    def _gather_ids(self, ids):
        ids = ids.reshape(-1).to(device=self.ctx.device, dtype=torch.int64).contiguous()
        if self.ctx.world > 1:
            import torch.distributed as dist
            n = torch.tensor([ids.numel()], dtype=torch.int64, device=self.ctx.device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.ctx.group)
            nmax = int(n[0])
            if ids.numel() < nmax:       # -1 = invalid id, skipped by the kernel
                ids = torch.cat([ids, torch.full((nmax - ids.numel(),), -1, dtype=torch.int64, device=ids.device)])
            out = torch.empty(self.ctx.world * nmax, dtype=torch.int64, device=ids.device)
            dist.all_gather_into_tensor(out, ids, group=self.ctx.group)
            ids = out
        return ids

    def _maybe_evict(self, incoming, stream):
        self._size_ub += incoming
        if self._size_ub <= self.max_load * self.capacity:
            return False
        size = self.eng.table_size(self.meta.handle)        # device sync; rare
        self._size_ub = size + incoming
        if self._size_ub <= self.max_load * self.capacity:
            return False
        target = int(self.evict_to * self.capacity)
        _native.cuda_check(self.lib.exb_tier_evict(self.h, target, self.clock, stream), "tier_evict")
        self._size_ub = target + incoming
        return True

    def prefetch(self, ids, ahead=False):
        """Make the rows of the batch resident. ``ids``: this rank's lookups (with world > 1 every rank's ids are
        gathered: a rank must hold the rows its PEERS are about to pull). ``ahead=True``: the ids belong to the
        NEXT batch -- the admit kernel runs on the tier's side stream while the current step computes (the
        pull-triggered asynchronous promotion of the reference). Rows are stamped with the tier's clock (one tick
        per batch admitted); eviction -- which moves slots -- only ever happens here, at the quiescent point
        before a batch's pull, never on the side stream."""
        key = (ids.data_ptr(), ids.numel())
        cur = torch.cuda.current_stream(self.ctx.device)
        if not ahead:
            self.clock += 1
            if self._prefetched == key:
                self._prefetched = None
                cur.wait_event(self._ev)
                return
        gathered = self._gather_ids(ids)
        n = gathered.numel()
        if n > self.max_load * self.capacity:
            uniq = int(torch.unique(gathered).numel())
            if uniq > self.max_load * self.capacity:
                from .status import Status, StatusError
                raise StatusError(Status.OOM, "host tier: one batch touches %d rows but the HBM cache holds %d "
                                  "(raise host_tier_rows / server.cache_size or pull in smaller pieces)"
                                  % (uniq, int(self.max_load * self.capacity)))
        if ahead:
            if self._size_ub + n > self.max_load * self.capacity:
                return          # would need an eviction: the regular call at the next step does it synchronously
            self._size_ub += n
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                _native.cuda_check(self.lib.exb_tier_admit(self.h, gathered.data_ptr(), n, self.clock + 1,
                                                           self.side.cuda_stream), "tier_admit")
                gathered.record_stream(self.side)
                self._ev.record(self.side)
            self._prefetched = key
            return
        self._maybe_evict(n, cur.cuda_stream)
        _native.cuda_check(self.lib.exb_tier_admit(self.h, gathered.data_ptr(), n, self.clock, cur.cuda_stream), "tier_admit")
        if self.ctx.world > 1:
            torch.cuda.synchronize(self.ctx.device)     # a peer may pull from this shard right after its own admit
            self.ctx.barrier()

    def mark_updated(self, ids):
        pass                      # dirtiness is implied by the stamps (stamp >= clean)

    def next_work(self):
        self.work_id += 1

    # ---- write-back
    def flush(self, before_work_id=None):
        st = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _native.cuda_check(self.lib.exb_tier_flush(self.h, self.clock, st), "tier_flush")
        torch.cuda.synchronize(self.ctx.device)

    def stats(self):
        out = (ctypes.c_uint64 * 8)()
        _native.cuda_check(self.lib.exb_tier_stats(self.h, out), "tier_stats")
        names = ["hits", "misses_host", "misses_new", "evicted", "writebacks", "host_rows", "resident", "host_capacity"]
        d = {k: int(out[i]) for i, k in enumerate(names)}
        look = d["hits"] + d["misses_host"] + d["misses_new"]
        d["miss_rate"] = (d["misses_host"] + d["misses_new"]) / look if look else 0.0
        return d

    def memory(self):
        out = (ctypes.c_uint64 * 5)()
        self.lib.exb_tier_info(self.h, out)
        return {"pinned_host_bytes": int(out[3]), "hbm_index_bytes": int(out[4]), "hbm_cache_bytes": self.cache_bytes}

    # ---- host store access (checkpoint side)
    def _slab(self):
        out = (ctypes.c_uint64 * 5)()
        self.lib.exb_tier_info(self.h, out)
        n = int(out[2]) * int(out[1])
        buf = (ctypes.c_float * n).from_address(int(out[0]))
        return np.ctypeslib.as_array(buf).reshape(int(out[2]), int(out[1]))

    def num_items(self):
        return self.stats()["host_rows"]

    def iter_rows(self, block_rows, with_state=True):
        """(local indices, weights, states) of every row of the host store, reference state layout"""
        cap = self._slab().shape[0]
        slots = torch.empty(cap, dtype=torch.int64, device=self.ctx.device)
        n = (ctypes.c_uint64 * 1)()
        _native.cuda_check(self.lib.exb_tier_host_enumerate(self.h, slots.data_ptr(), cap, n), "tier_enumerate")
        n = int(n[0])
        if n == 0:
            return
        from .ops.p2p_allreduce import tensor_from_ptr
        hk = tensor_from_ptr(self.lib.exb_tier_hkeys_ptr(self.h), cap * 2, self.ctx.device, dtype=torch.int32).view(torch.int64)
        slots = torch.sort(slots[:n])[0]
        ids = hk[slots].cpu().numpy().astype(np.uint64)
        slots = slots.cpu().numpy()
        slab = self._slab()
        dim, ws, m = self.meta.dim, self.wstride, self.meta
        from .config import OPTIMIZER_SLOTS
        nslots, nsc = OPTIMIZER_SLOTS(m.optimizer)
        for i in range(0, n, block_rows):
            sl = slots[i:i + block_rows]
            rows = slab[sl]
            w = np.ascontiguousarray(rows[:, :dim])
            if with_state and self.sd:
                parts = [rows[:, ws + s * ws: ws + s * ws + dim] for s in range(nslots)]
                parts.append(rows[:, ws + nslots * ws: ws + nslots * ws + nsc])
                st = np.ascontiguousarray(np.concatenate(parts, axis=1))
            else:
                st = np.empty((sl.size, 0), dtype=np.float32)
            yield (ids[i:i + block_rows] // np.uint64(m.shard_num)).astype(np.uint64), w, st

    def put_rows(self, global_ids, weights, states):
        """store the rows this rank owns in the host slab (load / restore); the cache is left alone"""
        m = self.meta
        ids = np.asarray(global_ids, dtype=np.uint64)
        owner = (m.shard_base + (ids % np.uint64(m.shard_num)).astype(np.int64)) % self.ctx.world
        sel = owner == self.ctx.rank
        if not sel.any():
            return
        ids = np.ascontiguousarray(ids[sel])
        w = np.asarray(weights, dtype=np.float32)[sel]
        from .config import OPTIMIZER_SLOTS
        nslots, nsc = OPTIMIZER_SLOTS(m.optimizer)
        rows = np.zeros((ids.size, self.HR), dtype=np.float32)
        rows[:, :m.dim] = w
        st = np.asarray(states, dtype=np.float32)
        ws = self.wstride
        if st.size and st.shape[1] == self.sd and self.sd:
            st = st[sel]
            for s in range(nslots):
                rows[:, ws + s * ws: ws + s * ws + m.dim] = st[:, s * m.dim:(s + 1) * m.dim]
            rows[:, ws + nslots * ws: ws + nslots * ws + nsc] = st[:, nslots * m.dim: nslots * m.dim + nsc]
        else:
            from .config import optimizer_slot_inits
            inits, scal = optimizer_slot_inits(m.optimizer)
            for s in range(nslots):
                rows[:, ws + s * ws: ws + s * ws + m.dim] = inits[s]
            for j in range(nsc):
                rows[:, ws + nslots * ws + j] = scal[j]
        rows = np.ascontiguousarray(rows)
        _native.cuda_check(self.lib.exb_tier_host_put(self.h, ids.ctypes.data, ids.size, rows.ctypes.data), "tier_host_put")

    def clear(self, host_too=True):
        _native.cuda_check(self.lib.exb_tier_clear(self.h, 1 if host_too else 0), "tier_clear")
        self._size_ub = 0
        self._prefetched = None

    # ---- lightweight checkpoint
    def persist(self, pool_dir, window=0):
        self.flush()
        os.makedirs(pool_dir, exist_ok=True)
        fn = os.path.join(pool_dir, "pool_v%d_r%d" % (self.meta.variable_id, self.ctx.rank))
        lib = _native.core()
        w = lib.exb_fw_open(fn.encode())
        n = self.num_items()
        lib.exb_fw_header(w, self.meta.variable_id, DTYPES[self.meta.dtype], self.meta.dim, self.meta.vocab, b"", 0,
                          0, 1, self.sd * 4, n)
        for idx, ww, ss in self.iter_rows(1 << 15):
            gid = np.ascontiguousarray(idx * np.uint64(self.meta.shard_num) +
                                       np.uint64((self.ctx.rank - self.meta.shard_base) % self.ctx.world))
            lib.exb_fw_block(w, gid.size, gid.ctypes.data, ww.ctypes.data, ww.nbytes, ss.ctypes.data if ss.nbytes else None,
                             ss.nbytes)
        lib.exb_fw_close(w)
        self.checkpoint = self.work_id
        self.pool_path = pool_dir
        return {"host_pool_path": pool_dir, "checkpoint": int(self.checkpoint)}

    def restore(self, pool_dir):
        from .checkpoint import iter_shard_file
        fn = os.path.join(pool_dir, "pool_v%d_r%d" % (self.meta.variable_id, self.ctx.rank))
        self.clear(host_too=True)
        if not os.path.exists(fn):
            return
        for rec in iter_shard_file(fn):
            if rec[0] != "block":
                continue
            _, hdr, gid, w, s = rec          # pool files store GLOBAL ids with shard_num 1 / shard_id 0
            self.put_rows(gid, w, s)
        self.pool_path = pool_dir

    @property
    def resident(self):          # should_persist() compatibility: {slot: work_id} is not materialised on the GPU
        return {}

    def close(self):
        if self.h:
            self.lib.exb_tier_destroy(self.h)
            self.h = None
            cache_budget(self.ctx).release(self.cache_bytes)
        _tiers.pop(self.meta.variable_id, None)


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

    # ---- checkpoint side: the store is authoritative once the dirty cache rows are written back
    def num_items(self):
        return int(self.lib.exb_var_num_items(self.store))

    def iter_rows(self, block_rows, with_state=True):
        cursor = ctypes.c_uint64(0)
        while True:
            idx = np.empty(block_rows, dtype=np.uint64)
            k = int(self.lib.exb_var_read_indices(self.store, ctypes.byref(cursor), idx.ctypes.data, block_rows))
            if k == 0:
                return
            idx = idx[:k]
            w = np.empty((k, self.dim), dtype=self.np_dt)
            s = np.empty((k, max(self.sd, 1)), dtype=self.np_dt)
            self.lib.exb_var_get_weights(self.store, idx.ctypes.data, k, w.ctypes.data, s.ctypes.data if self.sd else None)
            st = np.ascontiguousarray(s[:, :self.sd]) if (with_state and self.sd) else np.empty((k, 0), dtype=self.np_dt)
            yield (idx // np.uint64(self.meta.shard_num)).astype(np.uint64), w, st

    def put_rows(self, global_ids, weights, states):
        ids = np.asarray(global_ids, dtype=np.uint64)
        owner = (self.meta.shard_base + (ids % np.uint64(self.meta.shard_num)).astype(np.int64)) % self.ctx.world
        sel = owner == self.ctx.rank
        if not sel.any():
            return
        ids = np.ascontiguousarray(ids[sel])
        w = np.ascontiguousarray(np.asarray(weights)[sel], dtype=self.np_dt)
        st = np.asarray(states)
        has_state = st.size > 0 and st.shape[1] == self.sd and self.sd > 0
        s = np.ascontiguousarray(st[sel], dtype=self.np_dt) if has_state else None
        self.lib.exb_var_set_weights(self.store, ids.ctypes.data, ids.size, w.ctypes.data,
                                     s.ctypes.data if s is not None else None, self.sd * w.itemsize if s is not None else 0)

    def clear(self, host_too=True):
        with self._lock:
            self.be.clear(self.meta)
            self.resident = {}
            if host_too:
                self.lib.exb_var_clear(self.store)
                self._sync_store_config()

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
def make_tiered(ctx, meta, cache_rows=None, host_rows=None):
    """Attach a host tier to a (hash) variable. cache_rows defaults to the EnvConfig
    ``server.cache_size`` (MB) budget divided by the row footprint."""
    if cache_rows is None:
        mb = int(ctx.env["server"]["cache_size"])
        row_bytes = (meta.dim + optimizer_state_dim(meta.optimizer, meta.dim)) * 4 + 16
        cache_rows = max(1024, mb * (1 << 20) * 2 // 3 // row_bytes)
    root = ctx.env["server"].get("host_tier_root_path") or ctx.env["server"].get("pmem_pool_root_path") or None
    if ctx.device.type == "cuda":
        return GpuTieredVariable(ctx, meta, cache_rows, host_rows=host_rows, pool_path=root)
    return TieredVariable(ctx, meta, cache_rows, pool_path=root)


def tier_of(meta):
    t = _tiers.get(meta.variable_id)
    return t if (t is not None and t.meta is meta) else None


def close_all():
    """drop every tier of the process context (Context.finalize)"""
    global _budget
    for t in list(_tiers.values()):
        try:
            t.close()
        except Exception:
            pass
    _tiers.clear()
    _budget = None


def should_persist(ctx):
    """True when some tier has accumulated enough dirty rows that a lightweight checkpoint is
    worthwhile (reference piggy-backs this on pull responses, EmbeddingPullOperator.cpp:184-189)."""
    for t in _tiers.values():
        if isinstance(t, GpuTieredVariable):
            if t.work_id - max(t.checkpoint, 1) >= 64:      # batches since the last lightweight checkpoint
                return True
            continue
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
