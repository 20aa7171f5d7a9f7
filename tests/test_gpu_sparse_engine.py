"""CUDA sparse engine vs. the CPU oracle (libexb_core), same math header, fp32.

Model: the reference's c_api_test `pull_push` / `trd` / `mix` (openembedding/entry/
c_api_test.h:189-355): 1..N nodes x {array, hash} x dims, duplicate keys, exact checks
against a host-side oracle. "N nodes" are virtual ranks here: N engines in one process on
one GPU whose peer pointers reference each other -- the multi-rank dispatch / barrier /
combine protocol runs for real, only NVLink is replaced by local HBM.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle(dim, vocab, is_hash, init_cfg, opt_cfg, variable_id):
    from openembedding_b200 import _native
    from openembedding_b200.config import initializer_params, mix_seed, optimizer_params
    lib = _native.core()
    h = lib.exb_var_create(0x104, dim, 0 if is_hash else vocab, 0, 1, 1 if is_hash else 0)
    kind, p, seed = initializer_params(init_cfg)
    lib.exb_var_set_initializer(h, kind, p[0], p[1], p[2], mix_seed(seed, variable_id))
    ok, op = optimizer_params(opt_cfg)
    lib.exb_var_set_optimizer(h, ok, (ctypes.c_double * 8)(*op), 8)
    return lib, h


def _oracle_pull(lib, h, ids, dim):
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    out = np.empty((ids.size, dim), dtype=np.float32)
    lib.exb_var_pull(h, ids.ctypes.data, ids.size, out.ctypes.data)
    return out


def _run(world, specs, batch, steps, opt_cfg, init_cfg, seed=0, mode="stateless"):
    """specs: list of (dim, vocab, is_hash). Returns max abs error vs oracle after `steps`.
    mode: "stateless" -- pull(train=False) + push_update (the push plans the batch itself);
          "train"     -- pull(train=True): planned batch, unique remote rows (exb_pull2_kernel at world > 1);
          "prefetch"  -- like train, and the plan of step s+1 is built with prepare(next=True) during step s."""
    from openembedding_b200.ops.sparse_engine import CudaEngine
    torch.manual_seed(seed)
    dev = torch.device("cuda", 0)
    engines = [CudaEngine(0, r, world, max_ctas=6) for r in range(world)]
    F = len(specs)
    for e in engines:
        for vid, (dim, vocab, is_hash) in enumerate(specs):
            t = e.add_table(dim, vocab, is_hash, capacity=1 << 14)
            e.set_initializer(t, init_cfg, vid)
            e.set_optimizer(t, opt_cfg)
            e.alloc(t)
    plans = [e.make_plan(list(range(F)), batch) for e in engines]
    if world > 1:
        CudaEngine.connect_local(engines)
    oracles = [_oracle(dim, vocab, is_hash, init_cfg, opt_cfg, vid) for vid, (dim, vocab, is_hash) in enumerate(specs)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    sl = plans[0].feature_slices()
    worst = 0.0
    all_ids = []
    for step in range(steps):
        row = []
        for r in range(world):
            cols = []
            for (dim, vocab, is_hash) in specs:
                hi = min(vocab, 500) if not is_hash else 500
                c = torch.randint(0, hi, (batch,), dtype=torch.int64)
                if is_hash:
                    c = c * 1000003 + 7          # sparse keys in a huge space
                cols.append(c)
            row.append(torch.stack(cols, dim=1).contiguous().to(dev))
        all_ids.append(row)
    for step in range(steps):
        ids, grads, outs = all_ids[step], [], []
        for r in range(world):
            grads.append(torch.randn(batch, plans[r].io_stride, device=dev))
        torch.cuda.synchronize()
        # --- pull on every rank and compare with the oracle BEFORE the update
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                outs.append(plans[r].pull(ids[r], train=(mode != "stateless")))
                if mode == "prefetch" and step + 1 < steps:
                    plans[r].prepare(all_ids[step + 1][r], next=True)
        torch.cuda.synchronize()
        for r in range(world):
            for f, (dim, vocab, is_hash) in enumerate(specs):
                lib, h = oracles[f]
                want = _oracle_pull(lib, h, ids[r][:, f].cpu().numpy(), dim)
                got = outs[r][:, sl[f]].cpu().numpy()
                worst = max(worst, float(np.abs(want - got).max()))
        # --- fused push+update on all ranks concurrently (they barrier with each other)
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                plans[r].push_update(ids[r], grads[r])
        torch.cuda.synchronize()
        for e in engines:
            e.check()
        for f, (dim, vocab, is_hash) in enumerate(specs):
            lib, h = oracles[f]
            for r in range(world):
                k = np.ascontiguousarray(ids[r][:, f].cpu().numpy(), dtype=np.uint64)
                g = np.ascontiguousarray(grads[r][:, sl[f]].cpu().numpy(), dtype=np.float32)
                lib.exb_var_push(h, k.ctypes.data, k.size, g.ctypes.data, None)
            lib.exb_var_update(h)
    # final comparison over the whole touched id range
    for f, (dim, vocab, is_hash) in enumerate(specs):
        lib, h = oracles[f]
        probe = torch.arange(0, batch, dtype=torch.int64)
        probe = probe % (min(vocab, 500) if not is_hash else 500)
        if is_hash:
            probe = probe * 1000003 + 7
        full = torch.zeros((batch, F), dtype=torch.int64)
        full[:, f] = probe
        got = plans[0].pull(full.to(dev))[:, sl[f]].cpu().numpy()
        want = _oracle_pull(lib, h, probe.numpy(), dim)
        worst = max(worst, float(np.abs(want - got).max()))
    for lib, h in oracles:
        lib.exb_var_destroy(h)
    for e in engines:
        e.check()
        e.close()
    return worst


ADAGRAD = {"category": "adagrad", "learning_rate": 0.1}
UNIFORM = {"category": "uniform", "minval": -0.5, "maxval": 0.5}


@pytest.mark.parametrize("dim", [1, 3, 8, 9, 16, 64, 128, 200])
@pytest.mark.parametrize("is_hash", [False, True])
def test_single_rank_dims(dim, is_hash):
    err = _run(1, [(dim, 1000, is_hash)], batch=257, steps=3, opt_cfg=ADAGRAD, init_cfg=UNIFORM)
    assert err < 2e-4, err


@pytest.mark.parametrize("cat", ["default", "adadelta", "adagrad", "adam", "adamax", "ftrl", "rmsprop", "sgd", "test"])
def test_single_rank_optimizers(cat):
    cfg = {"category": cat, "learning_rate": 0.05}
    if cat == "sgd":
        cfg["momentum"] = 0.9
    err = _run(1, [(16, 300, False), (8, 0, True)], batch=128, steps=4, opt_cfg=cfg, init_cfg=UNIFORM)
    assert err < (5e-3 if cat == "test" else 5e-4), (cat, err)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_virtual_ranks(world):
    specs = [(64, 100000, False), (1, 100000, False), (9, 3, False), (16, 0, True), (4, 77, False)]
    err = _run(world, specs, batch=192, steps=3, opt_cfg=ADAGRAD, init_cfg=UNIFORM)
    assert err < 5e-4, err


@pytest.mark.parametrize("world,mode", [(1, "train"), (1, "prefetch"), (2, "train"), (3, "prefetch"), (4, "train"),
                                        (8, "prefetch")])
def test_planned_batches(world, mode, monkeypatch):
    """v2 path end to end: de-duplicated plan, unique-row pull (exb_pull2_kernel), pre-reduced push, prefetch"""
    monkeypatch.setenv("EXB_SPARSE_V2", "1")        # the default is v2 only on one GPU
    specs = [(64, 100000, False), (1, 100000, False), (9, 3, False), (16, 0, True), (4, 77, False), (200, 50, False)]
    err = _run(world, specs, batch=192, steps=4, opt_cfg=ADAGRAD, init_cfg=UNIFORM, mode=mode)
    assert err < 5e-4, err


def test_planned_test_optimizer_counts(monkeypatch):
    """the `test` optimizer consumes the per-id counts: they must survive de-duplication and pre-reduction"""
    monkeypatch.setenv("EXB_SPARSE_V2", "1")
    cfg = {"category": "test", "learning_rate": 0.05}
    err = _run(2, [(16, 300, False), (8, 0, True)], batch=128, steps=4, opt_cfg=cfg, init_cfg=UNIFORM, mode="train")
    assert err < 5e-3, err


def test_v1_kernels_still_match(monkeypatch):
    monkeypatch.setenv("EXB_SPARSE_V2", "0")
    specs = [(64, 100000, False), (1, 100000, False), (16, 0, True)]
    err = _run(2, specs, batch=192, steps=3, opt_cfg=ADAGRAD, init_cfg=UNIFORM)
    assert err < 5e-4, err


def test_abandoned_plan_is_reset():
    """a planned batch that is never pushed (evaluation under grad mode) must not leak into the next batch"""
    from openembedding_b200.ops.sparse_engine import CudaEngine
    dev = torch.device("cuda", 0)
    e = CudaEngine(0, 0, 1)
    t = e.add_table(8, 1000, False)
    e.set_initializer(t, {"category": "constant", "value": 0.0}, 0)
    e.set_optimizer(t, {"category": "sgd", "learning_rate": 1.0})
    e.alloc(t)
    plan = e.make_plan([t], 256)
    a = torch.arange(0, 256, dtype=torch.int64, device=dev).reshape(-1, 1) % 7
    b = (torch.arange(0, 256, dtype=torch.int64, device=dev).reshape(-1, 1) % 5) + 100
    plan.pull(a, train=True)                  # planned, never pushed
    plan.prepare(b, next=True)                # prefetched, never used
    g = torch.ones(256, plan.io_stride, device=dev)
    plan.pull(b, train=True)
    plan.push_update(b, g)                    # must update ids 100..104 only
    plan.push_update(a, g)
    torch.cuda.synchronize()
    e.check()
    probe = torch.tensor([[0], [6], [7], [100], [104], [105]], dtype=torch.int64, device=dev)
    out = plan.pull(probe)[:, 0].cpu()
    want = torch.tensor([-37.0, -36.0, 0.0, -52.0, -51.0, 0.0])    # 256 lookups over 7 / 5 ids, sgd lr 1
    assert torch.equal(out, want), out
    st = e.status()[1]
    assert st["pull_unique"] == 12, st
    e.close()


def test_hot_rows_many_duplicates():
    # every id identical: one row receives batch*world gradients (atomic contention path)
    from openembedding_b200.ops.sparse_engine import CudaEngine
    dev = torch.device("cuda", 0)
    e = CudaEngine(0, 0, 1)
    t = e.add_table(32, 10, False)
    e.set_initializer(t, {"category": "constant", "value": 1.0}, 0)
    e.set_optimizer(t, {"category": "sgd", "learning_rate": 1.0})
    e.alloc(t)
    plan = e.make_plan([t], 4096)
    ids = torch.full((4096, 1), 3, dtype=torch.int64, device=dev)
    g = torch.full((4096, plan.io_stride), 0.5, device=dev)
    plan.push_update(ids, g)
    out = plan.pull(ids[:4])
    torch.cuda.synchronize()
    e.check()
    assert torch.allclose(out, torch.full_like(out, 1.0 - 2048.0))
    code, stats = e.status()
    assert stats["update_unique"] == 1 and stats["push_indices"] == 4096
    e.close()


def test_checkpoint_row_access_and_rehash():
    from openembedding_b200.ops.sparse_engine import CudaEngine
    dev = torch.device("cuda", 0)
    e = CudaEngine(0, 0, 1)
    t = e.add_table(8, 0, True, capacity=1024)
    e.set_initializer(t, UNIFORM, 0)
    e.set_optimizer(t, {"category": "adam", "learning_rate": 0.1})
    e.alloc(t)
    plan = e.make_plan([t], 512)
    ids = (torch.arange(400, dtype=torch.int64) * 7919 + 1).reshape(-1, 1).to(dev)
    g = torch.randn(400, plan.io_stride, device=dev)
    plan.push_update(ids, g)
    torch.cuda.synchronize()
    e.check()
    assert e.table_size(t) == 400
    keys = e.enumerate_ids(t)
    assert keys.numel() == 400 and torch.equal(keys, torch.sort(ids.reshape(-1))[0])
    w, s = e.gather_rows(t, keys)
    assert s.shape[1] == 2 * 8 + 2
    e.rehash(t, 4096)
    e.commit()
    w2, s2 = e.gather_rows(t, keys)
    assert torch.equal(w, w2) and torch.equal(s, s2)
    out = plan.pull(ids)
    order = torch.argsort(ids.reshape(-1))
    assert torch.equal(out[order][:, :8], w)
    e.clear_table(t)
    assert e.table_size(t) == 0
    e.scatter_rows(t, keys, w, s)
    w3, s3 = e.gather_rows(t, keys)
    assert torch.equal(w, w3) and torch.equal(s, s3)
    e.close()


def test_context_version_guard():
    """a rank that moves a table slab (rehash) announces a new context version; the peers' kernels report
    SERVER_TOO_OLD_CTX until everybody has reconnected (reference: pico-ps ctx version / Status::SERVER_TOO_OLD_CTX)"""
    from openembedding_b200.ops.sparse_engine import CudaEngine
    from openembedding_b200.status import Status, StatusError
    dev = torch.device("cuda", 0)
    engines = [CudaEngine(0, r, 2, max_ctas=6) for r in range(2)]
    for e in engines:
        t = e.add_table(8, 0, True, capacity=1024)
        e.set_initializer(t, UNIFORM, 0)
        e.set_optimizer(t, ADAGRAD)
        e.alloc(t)
    plans = [e.make_plan([0], 128) for e in engines]
    CudaEngine.connect_local(engines)
    # even ids: owned by rank 0, so the pull below never dereferences rank 1's (freed) slabs
    ids = (torch.arange(128, dtype=torch.int64) * 7919 * 2).reshape(-1, 1).to(dev)
    before = plans[0].pull(ids).clone()
    torch.cuda.synchronize()
    engines[0].check()
    v0 = engines[1].lib.exb_engine_ctx_version(engines[1].h)
    engines[1].rehash(0, 4096)                      # rank 1 moves its slabs; rank 0 still maps the old ones
    assert engines[1].lib.exb_engine_ctx_version(engines[1].h) == v0 + 1
    plans[0].pull(ids)
    torch.cuda.synchronize()
    with pytest.raises(StatusError) as ei:
        engines[0].check()
    assert ei.value.status == Status.SERVER_TOO_OLD_CTX and ei.value.status.retryable
    CudaEngine.connect_local(engines)               # the refresh: re-map, accept the new versions
    after = plans[0].pull(ids)
    torch.cuda.synchronize()
    engines[0].check()
    engines[1].check()
    assert torch.equal(before, after)
    for e in engines:
        e.close()


@pytest.mark.parametrize("world,dim,v2", [(1, 10, 1), (1, 65, 1), (2, 65, 1), (2, 17, 1), (1, 65, 0), (2, 65, 0), (2, 10, 0)])
def test_split_row_feature(world, dim, v2, monkeypatch):
    """one table row [embedding(D) | linear(1)] feeding two places of the activation / gradient row"""
    monkeypatch.setenv("EXB_SPARSE_V2", str(v2))
    from openembedding_b200.ops.sparse_engine import CudaEngine
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    D = dim - 1
    Dp = (D + 3) // 4 * 4
    io = Dp * 2 + 8                      # two features' embedding blocks, then the linear columns
    lin0 = Dp * 2 + 1
    engines = [CudaEngine(0, r, world, max_ctas=6) for r in range(world)]
    for e in engines:
        for vid in range(2):
            t = e.add_table(dim, 5000, vid == 1, capacity=1 << 12)
            e.set_initializer(t, UNIFORM, vid)
            e.set_optimizer(t, ADAGRAD)
            e.alloc(t)
    plans = [e.make_plan([0, 1], 160, feat_offsets=[0, Dp], io_stride=io, feat_offsets2=[lin0, lin0 + 1],
                         feat_split=[D, D]) for e in engines]
    if world > 1:
        CudaEngine.connect_local(engines)
    oracles = [_oracle(dim, 5000, vid == 1, UNIFORM, ADAGRAD, vid) for vid in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    worst = 0.0
    for step in range(3):
        ids = [torch.stack([torch.randint(0, 300, (160,)), torch.randint(0, 300, (160,)) * 1000003 + 7], 1).contiguous().to(dev)
               for _ in range(world)]
        grads = [torch.randn(160, io, device=dev) for _ in range(world)]
        outs = []
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                outs.append(plans[r].pull(ids[r], train=True))
        torch.cuda.synchronize()
        for r in range(world):
            for f in range(2):
                lib, h = oracles[f]
                want = _oracle_pull(lib, h, ids[r][:, f].cpu().numpy(), dim)
                got = torch.cat([outs[r][:, f * Dp:f * Dp + D], outs[r][:, lin0 + f:lin0 + f + 1]], 1).cpu().numpy()
                worst = max(worst, float(np.abs(want - got).max()))
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                plans[r].push_update(ids[r], grads[r])
        torch.cuda.synchronize()
        for e in engines:
            e.check()
        for f in range(2):
            lib, h = oracles[f]
            for r in range(world):
                k = np.ascontiguousarray(ids[r][:, f].cpu().numpy(), dtype=np.uint64)
                g = torch.cat([grads[r][:, f * Dp:f * Dp + D], grads[r][:, lin0 + f:lin0 + f + 1]], 1)
                g = np.ascontiguousarray(g.cpu().numpy(), dtype=np.float32)
                lib.exb_var_push(h, k.ctypes.data, k.size, g.ctypes.data, None)
            lib.exb_var_update(h)
    probe = torch.stack([torch.arange(160) % 300, (torch.arange(160) % 300) * 1000003 + 7], 1).contiguous().to(dev)
    out = plans[0].pull(probe)
    for f in range(2):
        lib, h = oracles[f]
        want = _oracle_pull(lib, h, probe[:, f].cpu().numpy(), dim)
        got = torch.cat([out[:, f * Dp:f * Dp + D], out[:, lin0 + f:lin0 + f + 1]], 1).cpu().numpy()
        worst = max(worst, float(np.abs(want - got).max()))
    for lib, h in oracles:
        lib.exb_var_destroy(h)
    for e in engines:
        e.close()
    assert worst < 5e-4, worst
