"""Pull + push/update of the flagship embedding plan alone (no dense model): the target of
`ncu --set full --import-source on -k regex:exb_push_update -s 3 -c 1` captures, and a quick
event-timed check of the two sparse kernels with the in-kernel phase clock.

    python tools/sparse_probe.py [--vocab kaggle|1tb] [--iters 6] [--optimizer adagrad]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openembedding_b200 as oe  # noqa: E402
from openembedding_b200.context import get_context  # noqa: E402
from openembedding_b200.models.ctr import CRITEO_1TB_VOCAB_20M, CRITEO_KAGGLE_VOCAB  # noqa: E402
from openembedding_b200.models.fused_dense import FusedCTR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vocab", default="kaggle")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--optimizer", default="adagrad")
ap.add_argument("--ids", default="loguniform", choices=["loguniform", "sequential"],
                help="sequential: consecutive rows per feature (DRAM-friendly) -- separates DRAM random-access cost from latency")
a = ap.parse_args()
oe.flags.device = "cuda"
ctx = get_context()
vocab = CRITEO_KAGGLE_VOCAB if a.vocab == "kaggle" else CRITEO_1TB_VOCAB_20M
m = FusedCTR(vocab, embedding_dim=64, model="deepfm", batch=a.batch, cache_threshold=a.batch,
             sparse_optimizer={"category": a.optimizer})
g = m.group
dev = ctx.device
gen = torch.Generator().manual_seed(0)
v = torch.tensor(vocab, dtype=torch.float64)
for it in range(a.iters):
    u = torch.rand((a.batch, 26), generator=gen, dtype=torch.float64)
    ids = (torch.floor(torch.exp(u * torch.log(v))) - 1).clamp_(min=0).to(torch.int64)
    if a.ids == "sequential":
        base = torch.randint(0, 1 << 20, (26,), generator=gen)
        ids = ((base[None, :] + torch.arange(a.batch)[:, None]) % torch.tensor(vocab)[None, :]).to(torch.int64)
    ids = ids.contiguous().to(dev)
    m.G32.normal_()
    torch.cuda.synchronize()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    g.pull(ids, out=m.X32)
    e1.record()
    g.push_update(ids, m.G32)
    e2.record()
    torch.cuda.synchronize()
    st = ctx.backend.engine.status()[1]
    print("iter %d pull %.1f us push+update %.1f us phases %s unique_rows %s" % (
        it, e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3, st["last_push_update_us"], st.get("update_unique")), flush=True)

# ---- per-warp trace of the apply phase (last iteration repeated with tracing on)
tr = g.enable_trace()
if tr is not None:
    g.pull(ids, out=m.X32)
    g.push_update(ids, m.G32)
    torch.cuda.synchronize()
    t = tr.cpu()
    t0 = t[:, 0]
    act = t0 > 0
    base = int(t0[act].min())
    ntab64 = m.ns
    names = ["count", "ulist", "cmap", "row/touched", "gather", "math+store"]
    seg = {k: [] for k in names}
    n_heavy, n_light, ends = 0, 0, []
    for w in range(t.shape[0]):
        if not act[w]:
            continue
        last = int(t0[w])
        for k in range(1, g.TRACE_SLOTS - 7, 8):
            e = [int(x) for x in t[w, k:k + 7]]
            if e[1] == 0:
                break
            heavy = (e[0] >> 32) < ntab64
            n_heavy += heavy
            n_light += not heavy
            if heavy and e[6]:
                prev = last
                for nm, tt in zip(names, e[1:7]):
                    seg[nm].append((tt - prev) / 1e3)
                    prev = tt
            last = e[6] or last
        ends.append((last - base) / 1e3)
    import statistics as S
    q = lambda x: "mean %.2f p50 %.2f max %.2f" % (S.mean(x), S.median(x), max(x)) if x else "-"
    print("apply trace: warps %d heavy tasks %d light tasks %d (first %d tasks per warp traced)" % (
        int(act.sum()), n_heavy, n_light, (g.TRACE_SLOTS - 1) // 8))
    for nm in names:
        print("  heavy task %-12s us: %s" % (nm, q(seg[nm])))
    print("  warp finish after phase start (us): %s ; start skew %.2f" % (q(ends), (int(t0[act].max()) - base) / 1e3))
    g.enable_trace(False)
