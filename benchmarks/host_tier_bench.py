#!/usr/bin/env python
"""Host-DRAM tier benchmark (BASELINE.json config 4: "1B-row tables with host-DRAM overflow tier").

A hashed embedding table over a 1e9-id vocabulary is trained with a log-uniform (Zipf-like) id stream: the rows
that have ever been touched live in pinned host DRAM, an HBM cache of ``--cache-rows`` slots holds the hot ones.
Reports step time of the sparse path (admit + pull + push/update) tiered vs all-HBM and the cache miss rate --
the quantities of the reference's PMem evaluation (ICDE'23 paper SVI-C: miss rate 13.6 % with a 2 GB cache;
PMem-OE within 1-9 % of DRAM-PS).

    python benchmarks/host_tier_bench.py --steps 300 --cache-rows 1048576
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--features", type=int, default=8, help="lookups per sample (all into the one table)")
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--vocab", type=float, default=1e9)
    ap.add_argument("--cache-rows", type=int, default=1 << 20)
    ap.add_argument("--host-rows", type=int, default=1 << 24)
    ap.add_argument("--ahead", type=int, default=1, help="promote the next batch on the side stream")
    a = ap.parse_args()
    import openembedding_b200 as oe
    from openembedding_b200.context import get_context, reset_context
    from openembedding_b200.host_tier import make_tiered
    oe.flags.device = "cuda"
    oe.flags.config = "server:\n  cache_size: 65536\n"
    res = {}
    for mode in ("hbm", "tiered"):
        reset_context()
        ctx = get_context()
        dev = ctx.device
        st = ctx.create_storage(None)
        cap = 2 * a.cache_rows if mode == "tiered" else 4 * a.host_rows
        m = ctx.create_variable(st, 2 ** 63, a.dim, "float32", capacity=cap)
        ctx.set_initializer(m, {"category": "uniform", "minval": -0.05, "maxval": 0.05})
        ctx.set_optimizer(m, {"category": "adagrad", "learning_rate": 0.01})
        ctx.backend.ensure_allocated([m])
        tier = make_tiered(ctx, m, a.cache_rows, host_rows=a.host_rows) if mode == "tiered" else None
        F = a.features
        plan = ctx.backend.engine.make_plan([m.handle] * F, a.batch, feat_cols=list(range(F)), ncols=F)
        ctx.backend.engine.connect(None)
        g = torch.Generator().manual_seed(7)
        n = a.steps + a.warmup + 1
        u = torch.rand((n, a.batch, F), generator=g, dtype=torch.float64)
        ids = torch.floor(torch.exp(u * torch.log(torch.tensor(a.vocab, dtype=torch.float64)))).to(torch.int64)
        ids = ((ids * 2654435761) % int(a.vocab)).to(dev)
        grads = torch.randn(a.batch, plan.io_stride, device=dev) * 0.01
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(n - 1):
            if k == a.warmup:
                torch.cuda.synchronize()
                if tier is not None:
                    s0 = tier.stats()
                ev0.record()
            if tier is not None:
                tier.prefetch(ids[k].reshape(-1))
                if a.ahead:
                    tier.prefetch(ids[k + 1].reshape(-1), ahead=True)
            plan.pull(ids[k], train=True)
            plan.push_update(ids[k], grads)
        ev1.record()
        torch.cuda.synchronize()
        ctx.backend.engine.check()
        r = {"us_per_step": ev0.elapsed_time(ev1) * 1e3 / a.steps}
        if tier is not None:
            s1 = tier.stats()
            look = sum(s1[k] - s0[k] for k in ("hits", "misses_host", "misses_new"))
            r.update({"miss_rate": (s1["misses_host"] - s0["misses_host"] + s1["misses_new"] - s0["misses_new"]) / max(look, 1),
                      "miss_rate_host_only": (s1["misses_host"] - s0["misses_host"]) / max(look, 1),
                      "evicted": s1["evicted"], "writebacks": s1["writebacks"], "host_rows": s1["host_rows"],
                      "resident": s1["resident"], "memory": tier.memory()})
        res[mode] = r
    res["slowdown"] = res["tiered"]["us_per_step"] / res["hbm"]["us_per_step"]
    res["config"] = vars(a)
    print(json.dumps(res))
    reset_context()


if __name__ == "__main__":
    main()
