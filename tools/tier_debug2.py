"""diagnose tests/test_gpu_host_tier.py::test_tiered_equals_untiered_bitwise: untiered twice + tiered, row-level diffs"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import openembedding_b200 as oe
oe.flags.device = "cuda"
import openembedding_b200.torch as embed
from openembedding_b200 import host_tier
from openembedding_b200.context import get_context, reset_context
from test_gpu_host_tier import _train, _probe

def run(tier_rows, steps=40):
    reset_context(); host_tier._budget = None
    get_context()
    emb = embed.Embedding(-1, 8, embeddings_initializer="uniform", host_tier_rows=tier_rows, host_store_rows=1 << 14)
    opt = embed.distributed_optimizer(torch.optim.Adagrad(emb.parameters(), lr=0.1, initial_accumulator_value=0.1))
    losses = _train(emb, opt, steps, seed=1, vocab_hi=4000)
    probe = _probe(emb, torch.arange(4000) * 7919 + 3)
    get_context().backend.engine.check()
    st = emb.variable.tier.stats() if tier_rows else None
    return losses, probe, st

a = run(None); b = run(None); c = run(512); d = run(512)
for name, x, y in (("untiered vs untiered", a, b), ("untiered vs tiered", a, c), ("tiered vs tiered", c, d)):
    bad = (x[1] != y[1]).any(1).nonzero().reshape(-1)
    print(name, "losses equal", x[0] == y[0], "rows differing", bad.numel(), "max diff", float((x[1] - y[1]).abs().max()))
    if bad.numel():
        print("  first bad rows", bad[:8].tolist(), x[1][bad[0]].tolist(), y[1][bad[0]].tolist())
print("tier stats", c[2])
# how often did the training touch the bad rows?
g = torch.Generator().manual_seed(1)
cnt = torch.zeros(4000, dtype=torch.int64); trip = torch.zeros(4000, dtype=torch.int64)
for _ in range(40):
    x = torch.randint(0, 4000, (256,), generator=g); torch.rand(256, generator=g)
    bc = torch.bincount(x, minlength=4000); cnt += bc; trip += (bc >= 3).long()
bad = (a[1] != c[1]).any(1).nonzero().reshape(-1)
print("bad rows: touches", cnt[bad][:16].tolist(), "triple-dup steps", trip[bad][:16].tolist(), "| all rows with triple dups:", int((trip > 0).sum()))
