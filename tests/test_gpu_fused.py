"""Fused (all own kernels) DeepFM / WDL step vs a torch fp32 autograd reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(vocab, B, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in vocab], dim=1).contiguous().to(dev)
    dense = torch.rand(B, 13, generator=g).to(dev)
    labels = (torch.rand(B, generator=g) < 0.3).float().to(dev)
    return ids, dense, labels


@pytest.mark.parametrize("model,dim,cache", [("deepfm", 16, 64), ("deepfm", 64, 0), ("wdl", 8, 64), ("deepfm", 9, 64)])
def test_fused_step_matches_reference(cuda_context, model, dim, cache):
    from openembedding_b200.context import get_context
    from openembedding_b200.models.fused_dense import FusedCTR
    ctx = get_context()
    vocab = [1000, 50, 20000, 7, 3000] + [300] * 21
    B = 256
    m = FusedCTR(vocab, embedding_dim=dim, model=model, batch=B, cache_threshold=cache, lr=0.05,
                 sparse_optimizer={"category": "adagrad", "learning_rate": 0.05}, dw_splits=2)
    # give the embeddings non-trivial values first: a few training steps
    for s in range(3):
        m.forward_backward(*_batch(vocab, B, ctx.device, seed=s))
    ids, dense, labels = _batch(vocab, B, ctx.device, seed=99)
    loss = m.forward_backward(ids, dense, labels, update=False)
    torch.cuda.synchronize()
    ctx.backend.engine.check()
    ref_loss, g = m.reference(ids, dense, labels)
    assert abs(float(loss) - float(ref_loss)) < 5e-3, (float(loss), float(ref_loss))
    # dense parameter gradients
    for name in ["W0", "W1", "wout", "wd", "bias"]:
        o, n = m.segs[name]
        a, b = m.gtheta[o:o + n], g["theta"][o:o + n]
        err = float((a - b).abs().max())
        scale = float(b.abs().max()) + 1e-6
        assert err < 0.05 * scale + 2e-4, (name, err, scale)
    if m.nc:
        for name in ["cache_emb", "cache_lin"]:
            o, n = m.segs[name]
            a, b = m.gtheta[o:o + n], g["theta"][o:o + n]
            assert float((a - b).abs().max()) < 0.05 * float(b.abs().max()) + 2e-4, name
    # sparse-row gradients handed to push_update
    ge = m.G32[:, :m.ns * m.Dp]
    err = float((ge - g["emb"]).abs().max())
    assert err < 0.05 * float(g["emb"].abs().max()) + 2e-5, err
    gl = m.G32[:, m.lin0:m.lin0 + m.ns]
    assert torch.allclose(gl, g["lin"], atol=1e-6, rtol=1e-4)


def test_fused_trains_and_graph(cuda_context):
    from openembedding_b200.context import get_context
    from openembedding_b200.models.fused_dense import FusedCTR, FusedTrainer
    ctx = get_context()
    vocab = [1000, 50, 20000, 7, 3000] + [300] * 21
    B = 256
    m = FusedCTR(vocab, embedding_dim=16, model="deepfm", batch=B, cache_threshold=64, lr=0.05,
                 sparse_optimizer={"category": "adagrad", "learning_rate": 0.05})
    tr = FusedTrainer(m, use_graph=True)
    b = _batch(vocab, B, ctx.device)
    losses = [float(tr.step(*b)) for _ in range(10)]
    ctx.backend.engine.check()
    assert losses[-1] < losses[0] - 0.01, losses


@pytest.mark.parametrize("graph", [False, True])
def test_prefetch_next_batch_matches_plain(cuda_context, graph):
    """step(..., next_ids=) pulls the next batch's rows + plan beside this step's dense optimizer (the reference's
    pulling()); the training trajectory is the one of plain steps. Includes a break in the announced sequence."""
    from openembedding_b200.context import get_context, reset_context
    from openembedding_b200.models.fused_dense import FusedCTR, FusedTrainer
    vocab = [1000, 50, 20000, 7, 3000] + [300] * 21
    B = 256
    curves = []
    for prefetch in (False, True, "stable"):
        reset_context()
        ctx = get_context()
        m = FusedCTR(vocab, embedding_dim=16, model="deepfm", batch=B, cache_threshold=64, lr=0.05,
                     sparse_optimizer={"category": "adagrad", "learning_rate": 0.05})
        tr = FusedTrainer(m, use_graph=graph)
        batches = [_batch(vocab, B, ctx.device, seed=s) for s in range(4)]
        order = [0, 1, 2, 3, 0, 2, 1, 3, 3, 0]
        losses = []
        for k, i in enumerate(order):
            nxt = None
            if prefetch and k + 1 < len(order) and k != 4:        # k == 4: no announcement -> next step pulls up front
                nxt = batches[order[k + 1]][0]
            if prefetch and k == 6:                               # announce one batch, train another: plan is dropped
                nxt = batches[0][0]
            # "stable": the batches are resident tensors at fixed addresses -> graphs captured directly on them
            losses.append(float(tr.step(*batches[i], next_ids=nxt, stable=prefetch == "stable")))
        torch.cuda.synchronize()
        ctx.backend.engine.check()
        if prefetch == "stable" and graph:
            assert len(tr._stable) >= 3, tr._stable.keys()
        curves.append(losses)
    for a, b, c in zip(*curves):
        assert abs(a - b) < 2e-4 and abs(a - c) < 2e-4, curves


@pytest.mark.parametrize("cfg", [{"category": "adam", "learning_rate": 0.01},
                                 {"category": "ftrl", "learning_rate": 0.05, "l1_regularization_strength": 0.001},
                                 {"category": "adagrad", "learning_rate": 0.05}])
def test_fused_dense_optimizers_match_keras(cuda_context, cfg):
    """the dense optimizer of the fused step (exb_dense_opt_kernel: adagrad / adam / ftrl) vs the Keras formulas"""
    from test_optimizers import keras_reference
    from openembedding_b200.context import get_context
    from openembedding_b200.models.fused_dense import FusedCTR
    ctx = get_context()
    vocab = [1000, 50, 20000, 7, 3000] + [300] * 21
    B = 256
    m = FusedCTR(vocab, embedding_dim=8, model="deepfm", batch=B, cache_threshold=64,
                 sparse_optimizer={"category": "adagrad", "learning_rate": 0.05}, dense_optimizer=dict(cfg))
    theta0 = m.theta.detach().cpu().double().clone()
    grads = []
    for s in range(4):
        b = _batch(vocab, B, ctx.device, seed=s)
        m.forward_backward(*b, update=False)
        torch.cuda.synchronize()
        grads.append(m.gtheta.detach().cpu().double().clone())
        m.forward_backward(*b, update=True)
        torch.cuda.synchronize()
    ctx.backend.engine.check()
    ref = keras_reference(cfg, theta0.view(1, -1), [g.view(1, -1) for g in grads]).view(-1)
    got = m.theta.detach().cpu().double()
    err = float((got - ref).abs().max())
    moved = float((ref - theta0).abs().max())
    assert moved > 1e-4 and err < 2e-2 * moved + 1e-6, (cfg, err, moved)


@pytest.mark.parametrize("dense_opt", ["adagrad", "ftrl"])
def test_graph_trajectory_equals_eager_and_warmup_is_neutral(cuda_context, dense_opt):
    """the eager warm-up before the first graph capture (FusedCTR.warmup: every kernel of the step, zero-row push,
    dense optimizer on a restored snapshot) changes no parameter: same dense parameters and sparse rows after it,
    and the graph-driven loss curve is the eager one from the first step on"""
    from openembedding_b200.context import get_context, reset_context
    from openembedding_b200.models.fused_dense import FusedCTR, FusedTrainer
    vocab = [1000, 50, 20000, 7, 3000] + [300] * 21
    B = 256
    curves = []
    for graph in (False, True):
        reset_context()
        ctx = get_context()
        m = FusedCTR(vocab, embedding_dim=16, model="deepfm", batch=B, cache_threshold=64, lr=0.05,
                     sparse_optimizer={"category": "adagrad", "learning_rate": 0.05},
                     dense_optimizer={"category": dense_opt, "learning_rate": 0.05})
        batches = [_batch(vocab, B, ctx.device, seed=s) for s in range(3)]
        if graph:
            theta0, acc0 = m.theta.clone(), m.accum.clone()
            def rows():          # pad columns of the activation row are never written: start from zeros
                out = torch.zeros((B, m.group.io_stride), dtype=torch.float32, device=ctx.device)
                return m.group.pull(batches[0][0], out=out)
            rows0 = rows()
            m.warmup(*batches[0])
            torch.cuda.synchronize()
            assert torch.equal(m.theta, theta0) and torch.equal(m.accum, acc0)
            assert torch.equal(rows(), rows0)
            assert int(m.opt_step.item()) == 0
        tr = FusedTrainer(m, use_graph=graph)
        curves.append([float(tr.step(*batches[k % 3])) for k in range(7)])
        torch.cuda.synchronize()
        ctx.backend.engine.check()
    for a, b in zip(*curves):
        assert abs(a - b) < 2e-4, curves
