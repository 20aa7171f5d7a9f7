"""Server-model checkpoints in the reference's on-disk format (bit-compatible).

Layout (writer openembedding/client/Model.cpp:89-108 and
openembedding/server/EmbeddingDumpOperator.cpp:13-100, reader
openembedding/server/EmbeddingLoadOperator.cpp:58-111)::

    <path>/model_meta                      JSON, indent 4, {"model_sign","variables":[...],"version":"0.2"}
    <path>/<storage_id>/model_<node>_<file>   per (shard, variable): header + blocks
        header: u32 variable_id | i32 dtype | u64 dim | u64 vocab | u64 len + YAML config |
                i32 shard_id | i32 shard_num | u64 state_line_size | u64 num_items
        block : u64 n | u64 local_index[n] | T weights[n][dim] | byte states[n][state_line_size]
        (global id = local_index * shard_num + shard_id)

Loading re-shards: any rank count / shard count can read any checkpoint.

Differences from the reference: there are no server processes, so ``save_model`` is a
collective -- every rank streams its own HBM shard (device key compaction + row gather,
kernel K8) into ``model_<rank>_<file>`` with the native writer in ``libexb_core``.
"""
import ctypes
import json
import os

import numpy as np

from . import _native
from .config import DTYPE_NAMES, DTYPES, dump_variable_config, load_variable_config

FORMAT_VERSION = "0.2"


def block_rows(dim, itemsize, state_line_size):
    """rows per block, EmbeddingVariable.cpp:84-87"""
    return 1023 * 1024 // (dim * itemsize + state_line_size) + 1


def model_meta_dict(ctx):
    return {
        "model_sign": ctx.model_sign(),
        "variables": [{"datatype": m.dtype, "embedding_dim": m.dim, "vocabulary_size": m.vocab,
                       "storage_name": str(m.storage_id)} for m in ctx.variables],
        "version": FORMAT_VERSION,
    }


def read_model_meta(path):
    with open(os.path.join(path, "model_meta")) as fh:
        meta = json.load(fh)
    ver = meta.get("version", "unknown")
    if ver != FORMAT_VERSION:
        raise ValueError("OpenEmbedding model format version is %s, current version is %s." % (ver, FORMAT_VERSION))
    return meta


def save_model(ctx, path, include_optimizer=True, num_files=None, persist=None):
    """Collective. ``path`` may be a local directory, ``file://``, ``mem://null/`` (discard) or
    ``hdfs://`` (every rank stages its shard files locally and uploads them, utils/fs.py)."""
    from .utils.fs import Staging, URIConfig
    cfg = URIConfig(path)
    if cfg.is_null or cfg.is_local:
        return _save_model_local(ctx, path if cfg.is_null else cfg.path, include_optimizer, num_files, persist)
    with Staging(path, "w") as local:
        return _save_model_local(ctx, local, include_optimizer, num_files, persist)


def _save_model_local(ctx, path, include_optimizer=True, num_files=None, persist=None):
    """Collective. ``persist`` = dict(extra YAML keys) writes header-only records
    (num_items = 0), the lightweight host-tier checkpoint (EmbeddingDumpOperator.cpp:65-77)."""
    lib = _native.core()
    null_sink = path.startswith("mem://null/")
    num_files = num_files or int(ctx.env["server"]["server_dump_files"])
    if not null_sink and ctx.rank == 0:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "model_meta"), "w") as fh:
            fh.write(json.dumps(model_meta_dict(ctx), indent=4))
    ctx.barrier()
    be = ctx.backend
    be.synchronize()
    for st in ctx.storages:
        sdir = os.path.join(path, str(st.storage_id)) if not null_sink else path
        if not null_sink:
            os.makedirs(sdir, exist_ok=True)
        writers = {}
        for vid, meta in enumerate(st.variables):
            shard_id = be.shard_id(meta)
            if shard_id < 0:
                continue
            file_id = shard_id % num_files          # pico-ps operator/DumpOperator.h:73-76
            if file_id not in writers:
                fn = (sdir + "/model_%d_%d" % (ctx.rank, file_id)) if not null_sink else path
                w = lib.exb_fw_open(fn.encode())
                if not w:
                    raise IOError("cannot open " + fn)
                writers[file_id] = w
            w = writers[file_id]
            itemsize = 4 if meta.dtype == "float32" else 8
            sls = be.state_dim(meta) * itemsize if include_optimizer else 0
            # host-tier variables: the host store is authoritative once the dirty cache rows are written back
            # (the reference's PMem tables dump every row through the normal path too)
            from .host_tier import GpuTieredVariable, tier_of
            tier = tier_of(meta)
            gpu_tier = tier if isinstance(tier, GpuTieredVariable) else None
            if tier is not None and persist is None:
                tier.flush()
            if gpu_tier is not None:
                n_items = 0 if persist is not None else gpu_tier.num_items()
            else:
                n_items = 0 if persist is not None else (be.num_items(meta) if tier is None else tier.num_items())
            cfg = dump_variable_config(be.table_kind(meta), n_items if persist is None else (be.num_items(meta) if tier is None else tier.num_items()),
                                       meta.optimizer, meta.initializer,
                                       include_optimizer=include_optimizer, extra=persist).encode()
            lib.exb_fw_header(w, vid, DTYPES[meta.dtype], meta.dim, meta.vocab, cfg, len(cfg),
                              shard_id, meta.shard_num, sls, n_items)
            if n_items:
                written = 0
                rows_iter = (tier.iter_rows(block_rows(meta.dim, itemsize, sls), with_state=include_optimizer)
                             if tier is not None else
                             be.iter_local_rows(meta, block_rows(meta.dim, itemsize, sls), with_state=include_optimizer))
                for idx, wts, sts in rows_iter:
                    idx = np.ascontiguousarray(idx, dtype=np.uint64)
                    wts = np.ascontiguousarray(wts)
                    sts = np.ascontiguousarray(sts)
                    lib.exb_fw_block(w, idx.size, idx.ctypes.data, wts.ctypes.data, wts.nbytes,
                                     sts.ctypes.data if sts.nbytes else None, sts.nbytes if include_optimizer else 0)
                    written += idx.size
                if written != n_items:
                    raise RuntimeError("dump: enumerated %d rows, header promised %d" % (written, n_items))
        for w in writers.values():
            lib.exb_fw_close(w)
    ctx.barrier()


def iter_shard_file(path, want=None):
    """yield ('header', dict) then ('block', hdr, global ids, weights, states) records of one file. ``want(hdr)`` False:
    the segment's blocks are seeked past, not read (a rank that does not own the segment's shard)."""
    lib = _native.core()
    r = lib.exb_fr_open(path.encode())
    if not r:
        raise IOError("cannot open " + path)
    try:
        cfg = ctypes.create_string_buffer(1 << 20)
        while True:
            vid, dt = ctypes.c_uint32(), ctypes.c_int32()
            dim, vocab, clen = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
            sid, snum = ctypes.c_int32(), ctypes.c_int32()
            sls, items = ctypes.c_uint64(), ctypes.c_uint64()
            rc = lib.exb_fr_header(r, ctypes.byref(vid), ctypes.byref(dt), ctypes.byref(dim), ctypes.byref(vocab),
                                   cfg, len(cfg), ctypes.byref(clen), ctypes.byref(sid), ctypes.byref(snum),
                                   ctypes.byref(sls), ctypes.byref(items))
            if rc == 0:
                return
            if rc < 0:
                raise IOError("corrupt shard file " + path)
            hdr = {"variable_id": vid.value, "dtype": DTYPE_NAMES.get(dt.value, "unknown"), "dim": dim.value,
                   "vocab": vocab.value, "config": cfg.raw[:clen.value].decode(), "shard_id": sid.value,
                   "shard_num": snum.value, "state_line_size": sls.value, "num_items": items.value}
            yield ("header", hdr)
            np_dt = np.float32 if hdr["dtype"] == "float32" else np.float64
            itemsize = np.dtype(np_dt).itemsize
            done = 0
            skip = want is not None and not want(hdr)
            while done < hdr["num_items"]:
                n = lib.exb_fr_block_size(r)
                if n < 0:
                    raise IOError("truncated shard file " + path)
                if skip:
                    if lib.exb_fr_skip_block(r, n, n * hdr["dim"] * itemsize, n * hdr["state_line_size"]) != 0:
                        raise IOError("truncated shard file " + path)
                    done += n
                    continue
                idx = np.empty(n, dtype=np.uint64)
                w = np.empty((n, hdr["dim"]), dtype=np_dt)
                scols = hdr["state_line_size"] // itemsize
                s = np.empty((n, scols), dtype=np_dt)
                if lib.exb_fr_block(r, n, idx.ctypes.data, w.ctypes.data, w.nbytes,
                                    s.ctypes.data if s.nbytes else None, s.nbytes) != 0:
                    raise IOError("truncated shard file " + path)
                gid = idx * np.uint64(hdr["shard_num"]) + np.uint64(hdr["shard_id"])
                yield ("block", hdr, gid, w, s)
                done += n
    finally:
        lib.exb_fr_close(r)


def load_model(ctx, path, restore_config_only=False):
    from .utils.fs import Staging, URIConfig
    cfg = URIConfig(path)
    if cfg.is_local:
        return _load_model_local(ctx, cfg.path, restore_config_only)
    with Staging(path, "r") as local:
        return _load_model_local(ctx, local, restore_config_only)


def _load_model_local(ctx, path, restore_config_only=False):
    """Collective. Every rank walks the segment headers of every file, but reads only the segments it can own: a
    segment holds ONE saved shard (ids with ``id % shard_num == shard_id``), so with an unchanged ``shard_num`` its
    owner in the current layout is the single rank ``(shard_base + shard_id) % world`` -- whatever the world size was
    at save time -- and everybody else seeks past it. Only a changed ``shard_num`` (impossible today: the model meta
    must match) falls back to reading everything and filtering row by row (``load_rows``). ``restore_config_only``:
    initializer / optimizer configs are restored, no rows are loaded (reference: ``load_model`` with only the
    variable configs, used before a ``restore`` from the persistent tier)."""
    meta = read_model_meta(path)
    mine = model_meta_dict(ctx)["variables"]
    if meta["variables"] != mine:
        raise ValueError("model meta not match\n%s\n%s" % (json.dumps(meta["variables"], indent=4),
                                                           json.dumps(mine, indent=4)))
    be = ctx.backend
    from .host_tier import tier_of
    for m in ctx.variables:
        if tier_of(m) is not None:
            tier_of(m).clear(host_too=True)      # cache AND host store: the checkpoint replaces both
        else:
            be.clear(m)
    for st in ctx.storages:
        sdir = os.path.join(path, str(st.storage_id))
        if not os.path.isdir(sdir):
            continue
        def want(hdr, st=st):
            if restore_config_only:
                return False
            var = st.variables[hdr["variable_id"]]
            if int(hdr["shard_num"]) != int(var.shard_num):
                return True                      # re-sharded by id: every rank filters row by row
            return (int(var.shard_base) + int(hdr["shard_id"])) % ctx.world == ctx.rank

        for fn in sorted(os.listdir(sdir)):
            if not fn.startswith("model_"):
                continue
            for rec in iter_shard_file(os.path.join(sdir, fn), want):
                if rec[0] == "header":
                    hdr = rec[1]
                    var = st.variables[hdr["variable_id"]]
                    cfg = load_variable_config(hdr["config"])
                    # optimizer category change on load resets states (EmbeddingVariable.cpp:44-47)
                    if "initializer" in cfg:
                        ctx.set_initializer(var, cfg["initializer"])
                    if "optimizer" in cfg and cfg["optimizer"] != var.optimizer:
                        ctx.set_optimizer(var, cfg["optimizer"])
                    continue
                _, hdr, gid, w, s = rec
                var = st.variables[hdr["variable_id"]]
                if tier_of(var) is not None:
                    tier_of(var).put_rows(gid, w, s)     # rows land in the host store; the cache refills on demand
                else:
                    be.load_rows(var, gid, w, s)
    try:
        sign = meta.get("model_sign", "")
        ctx.loaded_model_sign = sign
    except Exception:
        pass
    be.synchronize()
    ctx.barrier()
