"""Host-DRAM overflow tier: a cache of N rows in front of a host store must behave exactly
like an untiered table (reference analogue: openembedding/entry/pmem_c_api_test.cpp --
persist / restore round trips, pending window)."""
import tempfile

import torch


def _train(emb, opt, steps, seed, vocab_hi):
    g = torch.Generator().manual_seed(seed)
    outs = []
    for _ in range(steps):
        x = torch.randint(0, vocab_hi, (64,), generator=g)
        y = torch.rand(64, generator=g)
        loss = ((emb(x).sum(-1) - y) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        outs.append(float(loss))
    return outs


def test_tiered_equals_untiered(cpu_context):
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context, reset_context
    res = []
    for tier_rows in (None, 40):       # 40-row cache in front of ~300 live rows: constant eviction
        reset_context()
        get_context()
        torch.manual_seed(0)
        emb = embed.Embedding(-1, 8, embeddings_initializer="uniform", host_tier_rows=tier_rows)
        opt = embed.distributed_optimizer(torch.optim.Adagrad(emb.parameters(), lr=0.1, initial_accumulator_value=0.1))
        losses = _train(emb, opt, 30, seed=1, vocab_hi=300)
        resident = len(emb.variable.tier.resident) if tier_rows else 0
        probe = emb(torch.arange(300)).detach().clone()
        res.append((losses, probe, emb, resident))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])
    t = res[1][2].variable.tier
    assert t.stats["flushes"] > 0 and t.stats["writebacks"] > 0 and res[1][3] <= 40 + 64


def test_persist_restore_roundtrip(cpu_context):
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context
    get_context()
    emb = embed.Embedding(-1, 4, embeddings_initializer="uniform", host_tier_rows=64)
    opt = embed.distributed_optimizer(torch.optim.Adam(emb.parameters(), lr=0.05))
    _train(emb, opt, 10, seed=2, vocab_hi=200)
    d = tempfile.mkdtemp()
    embed.persist_server_model(None, d + "/ck", 0)
    want = emb(torch.arange(200)).detach().clone()
    _train(emb, opt, 5, seed=3, vocab_hi=200)
    assert not torch.equal(emb(torch.arange(200)).detach(), want)
    embed.restore_server_model(None, d + "/ck")
    assert torch.equal(emb(torch.arange(200)).detach(), want)
    # optimizer state came back too: continuing from the restored state is reproducible
    a = _train(emb, opt, 3, seed=4, vocab_hi=200)
    embed.restore_server_model(None, d + "/ck")
    b = _train(emb, opt, 3, seed=4, vocab_hi=200)
    assert a == b
    assert isinstance(embed.should_persist_server_model(None), bool)


def test_tier_multi_rank_gloo():
    """world=2 over gloo: tiered == untiered, rank-local promotion / eviction / write-back"""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29700 + os.getpid() % 200
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, "mp_cpu_tier_check.py")],
                       env=dict(os.environ, OMP_NUM_THREADS="1"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300)
    assert r.returncode == 0 and "MP_CPU_TIER_CHECK_PASSED" in r.stdout, r.stdout[-3000:]
