#!/usr/bin/env python
"""Checkpoint throughput: dump (reference on-disk format, weights + optimizer state) and load of one rank's shard.

The reference publishes 78 GB in 869 s for Criteo-1TB (documents/en/benchmark.md:50-55) = 0.09 GB/s through its
CPU servers. Here: device key compaction -> gather kernel -> double-buffered pinned D2H -> native block writer.

    python benchmarks/checkpoint_bench.py --rows 8000000 --dim 64
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8_000_000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    import openembedding_b200 as oe
    from openembedding_b200 import checkpoint as ck
    from openembedding_b200.context import get_context
    oe.flags.device = "cuda"
    ctx = get_context()
    st = ctx.create_storage(None)
    m = ctx.create_variable(st, a.rows, a.dim, "float32")
    ctx.set_initializer(m, {"category": "uniform", "minval": -0.05, "maxval": 0.05})
    ctx.set_optimizer(m, {"category": "adagrad"})
    be = ctx.backend
    be.ensure_allocated([m])
    g = torch.Generator(device=ctx.device).manual_seed(0)
    for i in range(0, a.rows, 1 << 20):                     # materialise every row (array tables dump touched rows)
        ids = torch.arange(i, min(a.rows, i + (1 << 20)), device=ctx.device)
        be.engine.scatter_rows(m.handle, ids, torch.randn(ids.numel(), a.dim, device=ctx.device, generator=g),
                               torch.rand(ids.numel(), a.dim, device=ctx.device, generator=g))
    torch.cuda.synchronize()
    gb = a.rows * (8 + a.dim * 4 * 2) / 1e9
    res = {"rows": a.rows, "dim": a.dim, "gigabytes": round(gb, 3)}
    t0 = time.perf_counter()
    ck.save_model(ctx, "mem://null/", include_optimizer=True)
    res["dump_null_sink_s"] = time.perf_counter() - t0
    d = tempfile.mkdtemp(dir=a.dir if os.path.isdir(a.dir) else None)
    t0 = time.perf_counter()
    ck.save_model(ctx, d + "/model", include_optimizer=True)
    res["dump_file_s"] = time.perf_counter() - t0
    probe = torch.arange(0, a.rows, max(1, a.rows // 4096), device=ctx.device)[:4096]
    before, _ = be.engine.gather_rows(m.handle, probe)
    t0 = time.perf_counter()
    ck.load_model(ctx, d + "/model")
    torch.cuda.synchronize()
    res["load_file_s"] = time.perf_counter() - t0
    after, _ = be.engine.gather_rows(m.handle, probe)
    assert torch.equal(before, after)
    for k in ("dump_null_sink_s", "dump_file_s", "load_file_s"):
        res[k[:-2] + "_gbps"] = round(gb / res[k], 3)
    res["reference_gbps"] = round(78 / 869, 3)
    shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
