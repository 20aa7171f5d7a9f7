"""Host-DRAM tier on the CUDA engine (csrc/cuda/host_tier.cuh): pinned host slab + HBM cache/residency map.

Reference analogues: openembedding/variable/pmem_embedding_table_test.cpp (cache hit / miss / eviction),
openembedding/entry/pmem_c_api_test.cpp (persist / restore round trips)."""
import tempfile

import pytest
import torch


def get_ctx():
    from openembedding_b200.context import get_context
    return get_context()

pytestmark = pytest.mark.gpu


def _train(emb, opt, steps, seed, vocab_hi, n=256):
    g = torch.Generator().manual_seed(seed)
    outs = []
    for _ in range(steps):
        # every id at most twice per batch: the gradient rows are summed with float atomics, and only a two-term sum is
        # independent of their order (three duplicates of one id made one row differ by 1 ulp between identical runs)
        p = torch.randperm(vocab_hi, generator=g)[:n - n // 4]
        x = torch.cat([p, p[:n // 4]]) * 7919 + 3
        y = torch.rand(n, generator=g)
        out = emb(x)
        loss = ((out.sum(-1) - y.to(out.device)) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        outs.append(float(loss))
    return outs


def _probe(emb, ids, chunk=256):
    """read rows through the tier in cache-sized pieces"""
    with torch.no_grad():
        return torch.cat([emb(ids[i:i + chunk]).detach().cpu() for i in range(0, ids.numel(), chunk)])


def _reset():
    from openembedding_b200 import host_tier
    from openembedding_b200.context import get_context, reset_context
    reset_context()
    host_tier._budget = None
    return get_context()


def test_tiered_equals_untiered_bitwise(cuda_context):
    """a 1024-slot HBM cache in front of ~4000 live rows: evictions + write-backs + re-promotions, same bits"""
    import openembedding_b200.torch as embed
    res = []
    for tier_rows in (None, 512):
        _reset()
        emb = embed.Embedding(-1, 8, embeddings_initializer="uniform", host_tier_rows=tier_rows, host_store_rows=1 << 14)
        opt = embed.distributed_optimizer(torch.optim.Adagrad(emb.parameters(), lr=0.1, initial_accumulator_value=0.1))
        losses = _train(emb, opt, 40, seed=1, vocab_hi=4000)
        probe = _probe(emb, torch.arange(4000) * 7919 + 3)
        get_ctx().backend.engine.check()
        res.append((losses, probe, emb))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])
    st = res[1][2].variable.tier.stats()
    assert st["evicted"] > 0 and st["writebacks"] > 0 and st["misses_host"] > 0 and st["misses_new"] > 0, st
    assert st["resident"] <= 1024 and 0.0 < st["miss_rate"] < 1.0, st
    mem = res[1][2].variable.tier.memory()
    assert mem["pinned_host_bytes"] > 0 and mem["hbm_cache_bytes"] > 0


def test_prefetch_ahead_on_side_stream(cuda_context):
    """the next batch's rows are promoted on the tier's side stream while the current batch trains"""
    import openembedding_b200.torch as embed
    ctx = _reset()
    emb = embed.Embedding(-1, 16, embeddings_initializer="uniform", host_tier_rows=512, host_store_rows=1 << 14)
    ref_ctx_emb = None
    opt = embed.distributed_optimizer(torch.optim.SGD(emb.parameters(), lr=0.5))
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randint(0, 3000, (256,), generator=g) * 7919 + 3).to(ctx.device) for _ in range(20)]
    tier = emb.variable.tier
    for k, x in enumerate(batches):
        if k + 1 < len(batches):
            tier.prefetch(batches[k + 1], ahead=True)        # overlaps this step
        out = emb(x)
        loss = (out.sum(-1) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    ctx.backend.engine.check()
    got = _probe(emb, torch.arange(3000) * 7919 + 3)
    # same run without the tier
    ctx = _reset()
    emb2 = embed.Embedding(-1, 16, embeddings_initializer="uniform")
    opt2 = embed.distributed_optimizer(torch.optim.SGD(emb2.parameters(), lr=0.5))
    for x in batches:
        out = emb2(x.to(ctx.device))
        loss = (out.sum(-1) ** 2).mean()
        opt2.zero_grad()
        loss.backward()
        opt2.step()
    want = _probe(emb2, torch.arange(3000) * 7919 + 3)
    assert torch.equal(got, want)


def test_persist_restore_and_checkpoint(cuda_context):
    import openembedding_b200.torch as embed
    _reset()
    emb = embed.Embedding(-1, 4, embeddings_initializer="uniform", host_tier_rows=512, host_store_rows=1 << 13)
    opt = embed.distributed_optimizer(torch.optim.Adam(emb.parameters(), lr=0.05))
    ids = torch.arange(2500) * 7919 + 3
    _train(emb, opt, 12, seed=2, vocab_hi=2500)
    d = tempfile.mkdtemp()
    embed.persist_server_model(None, d + "/ck", 0)
    embed.save_server_model(None, d + "/full", include_optimizer=True)
    want = _probe(emb, ids)
    _train(emb, opt, 5, seed=3, vocab_hi=2500)
    assert not torch.equal(_probe(emb, ids), want)
    embed.restore_server_model(None, d + "/ck")
    assert torch.equal(_probe(emb, ids), want)
    a = _train(emb, opt, 3, seed=4, vocab_hi=2500)
    embed.restore_server_model(None, d + "/ck")
    b = _train(emb, opt, 3, seed=4, vocab_hi=2500)
    assert a == b                                           # optimizer state came back too
    # full checkpoint (reference format) written from the host store, loaded back through it
    embed.load_server_model(None, d + "/full")
    assert torch.equal(_probe(emb, ids), want)
    c = _train(emb, opt, 3, seed=4, vocab_hi=2500)
    assert a == c


def test_cache_budget_is_enforced(cuda_context):
    import openembedding_b200 as oe
    import openembedding_b200.torch as embed
    from openembedding_b200.status import StatusError
    old = oe.flags.config
    oe.flags.config = "server:\n  cache_size: 1\n"           # 1 MB of HBM cache in total
    try:
        _reset()
        with pytest.raises(StatusError):
            embed.Embedding(-1, 64, embeddings_initializer="zeros", host_tier_rows=1 << 16)
    finally:
        oe.flags.config = old
        _reset()
