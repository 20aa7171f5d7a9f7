"""Public API semantics on the CPU backend (mirrors the contract of openembedding/tensorflow/exb.py)."""
import os
import tempfile

import pytest
import torch


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.big = torch.nn.Embedding(1000, 8)
        self.small = torch.nn.Embedding(10, 4)
        self.out = torch.nn.Linear(12, 1)

    def forward(self, x, y):
        return self.out(torch.cat([self.big(x), self.small(y)], -1)).squeeze(-1)


def test_distributed_model_replaces_embeddings(cpu_context):
    import openembedding_b200.torch as embed
    m = embed.distributed_model(_Net(), sparse_as_dense_size=64)
    assert isinstance(m.big, embed.Embedding) and not m.big.sparse_as_dense
    assert isinstance(m.small, embed.Embedding) and m.small.sparse_as_dense
    assert m.big.embeddings.shape == (1, 8)            # dummy [1, dim] graph variable (exb.py:430-432)
    assert m.small.embeddings.shape == (10, 4)
    for name in ("save", "save_weights", "load_weights", "save_as_original_model"):
        assert hasattr(m, name)


def test_train_checkpoint_export(cpu_context):
    import openembedding_b200.torch as embed
    torch.manual_seed(0)
    m = embed.distributed_model(_Net())
    opt = embed.distributed_optimizer(torch.optim.Adagrad(m.parameters(), lr=0.1, initial_accumulator_value=0.1))
    x, y, t = torch.randint(0, 1000, (64,)), torch.randint(0, 10, (64,)), torch.rand(64)
    first = None
    for _ in range(20):
        loss = ((m(x, y) - t) ** 2).mean()
        opt.zero_grad(); loss.backward(); opt.step()
        first = first if first is not None else float(loss)
    assert float(loss) < first
    d = tempfile.mkdtemp()
    m.save_weights(d + "/w")
    assert os.path.exists(d + "/w.openembedding/openembedding/model_meta")
    snap = m.big(x).detach().clone()
    for _ in range(3):
        loss = ((m(x, y) - t) ** 2).mean()
        opt.zero_grad(); loss.backward(); opt.step()
    assert not torch.equal(m.big(x).detach(), snap)
    m.load_weights(d + "/w")
    assert torch.equal(m.big(x).detach(), snap)
    m.save(d + "/saved", include_optimizer=False)
    assert os.path.exists(d + "/saved/openembedding/model_meta") and os.path.exists(d + "/saved/model.pt")
    plain = m.save_as_original_model(d + "/plain.pt")
    assert type(plain.big) is torch.nn.Embedding and plain.big.weight.shape == (1000, 8)
    assert torch.allclose(plain.big.weight[x], snap)
    again = torch.load(d + "/plain.pt", weights_only=False)
    assert torch.allclose(again(x, y), m(x, y).detach().cpu(), atol=1e-6)
    with pytest.raises(ValueError):
        m.save_as_original_model(d + "/p2.pt", include_optimizer=True)


def test_hash_embedding_and_errors(cpu_context):
    import openembedding_b200.torch as embed
    h = embed.Embedding(-1, 4, embeddings_initializer="zeros")
    ids = torch.tensor([0, 2 ** 62, 12345678901234])
    assert torch.equal(h(ids), torch.zeros(3, 4))
    with pytest.raises(ValueError):
        embed.Embedding(100, 4, embeddings_regularizer=lambda w: w.sum())        # explicit=True rejects it
    embed.Embedding(100, 4, embeddings_regularizer=lambda w: w.sum(), explicit=False)
    with pytest.raises(ValueError):
        embed.Embedding(0, 4)
    p = [torch.nn.Parameter(torch.zeros(1))]
    with pytest.raises(ValueError):
        embed.distributed_optimizer(torch.optim.Adam(p, amsgrad=True))
    with pytest.raises(ValueError):
        embed.distributed_optimizer(torch.optim.RMSprop(p, centered=True))
    with pytest.raises(ValueError):
        embed.distributed_optimizer(torch.optim.NAdam(p))       # wrapped by the reference, no server impl
    with pytest.raises(ValueError):
        plain = embed.Embedding(-1, 4)
        wrap = torch.nn.Sequential(plain)
        embed.save_as_original_model(wrap, tempfile.mkdtemp() + "/x.pt")   # hash tables cannot be exported


def test_variable_verbs_and_sum_semantics(cpu_context):
    import openembedding_b200.torch as embed
    v = embed.Variable(initializer={"category": "constant", "value": 1.0}, shape=(50, 3))
    v.set_server_optimizer({"category": "sgd", "learning_rate": 0.5})
    idx = torch.tensor([3, 3, 7])
    fake = v.push_gradients(idx, torch.ones(3, 3))
    assert fake.shape == v.graph_var.shape
    v.update_weights(fake)
    out = v.sparse_read(torch.tensor([3, 7, 9]))
    assert torch.allclose(out[0], torch.full((3,), 0.0)) and torch.allclose(out[1], torch.full((3,), 0.5))
    assert torch.allclose(out[2], torch.ones(3))
    assert v.prefetch(idx) is not None
    dv = embed.distributed_variable(initializer="ones", shape=(5, 2), sparse_as_dense=True)
    assert dv.sparse_as_dense and dv.sparse_read(torch.tensor([1, 1])).shape == (2, 2)
    with pytest.raises(ValueError):
        dv.prefetch(idx)


def test_all_optimizer_classes(cpu_context):
    import openembedding_b200.torch as embed
    for cls, kw in [(embed.Adadelta, {}), (embed.Adagrad, {"lr": 0.1}), (embed.Adam, {}), (embed.Adamax, {}),
                    (embed.RMSprop, {}), (embed.SGD, {"lr": 0.1, "momentum": 0.9}), (embed.FtrlDistributed, {"lr": 0.1})]:
        e = embed.Embedding(100, 4)
        lin = torch.nn.Linear(4, 1)
        opt = cls(list(e.parameters()) + list(lin.parameters()), **kw)
        x = torch.randint(0, 100, (32,))
        l0 = None
        for _ in range(15):
            loss = (lin(e(x)).squeeze(-1) - 1.0).pow(2).mean()
            opt.zero_grad(); loss.backward(); opt.step()
            l0 = l0 if l0 is not None else float(loss)
        assert float(loss) < l0, cls.__name__
    with pytest.raises(ValueError):
        embed.Nadam([torch.nn.Parameter(torch.zeros(1))])


def test_pulling_prefetches_ids(cpu_context):
    import openembedding_b200.torch as embed
    e = embed.Embedding(100, 4, name="C1")
    m = torch.nn.Sequential(e)
    data = [({"C1": torch.randint(0, 100, (8,)), "I1": torch.rand(8)}, torch.rand(8)) for _ in range(5)]
    got = list(embed.pulling(data, m))
    assert len(got) == 5 and all(torch.equal(a[0]["C1"], b[0]["C1"]) for a, b in zip(got, data))
    assert len(list(embed.pulling(data, m, steps=3))) == 3


def test_embedding_multi_hot_bags(cpu_context):
    """ragged / multi-hot ids: flat values + offsets (EmbeddingBag semantics) and nested tensors"""
    import openembedding_b200.torch as embed
    emb = embed.Embedding(50, 4, embeddings_initializer={"category": "uniform", "minval": -1.0, "maxval": 1.0})
    flat = torch.tensor([3, 7, 7, 1, 9, 3])
    offsets = torch.tensor([0, 2, 2, 5])          # bags: [3,7] [] [7,1,9] [3]
    rows = emb(flat).detach()
    out = emb(flat, offsets=offsets)
    ref = torch.stack([rows[0:2].sum(0), torch.zeros(4), rows[2:5].sum(0), rows[5:6].sum(0)])
    assert torch.allclose(out, ref)
    assert torch.allclose(emb(flat, offsets=offsets, mode="mean")[2], rows[2:5].mean(0))
    assert torch.allclose(emb(flat, offsets=offsets, mode="max")[0], rows[0:2].max(0).values)
    w = torch.tensor([1.0, 2.0, 0.5, 1.0, 1.0, 3.0])
    assert torch.allclose(emb(flat, offsets=offsets, per_sample_weights=w)[3], 3.0 * rows[5])
    # gradients flow back to the rows of the bag (duplicates summed by the server-side reduce)
    opt = embed.distributed_optimizer(torch.optim.SGD(emb.parameters(), lr=1.0))
    before = emb(torch.tensor([7])).detach().clone()
    loss = emb(flat, offsets=offsets).sum()
    opt.zero_grad(); loss.backward(); opt.step()
    after = emb(torch.tensor([7])).detach()
    assert torch.allclose(after, before - 2.0)    # id 7 appears twice, d(loss)/d(row) = 1 each
    nt = torch.nested.nested_tensor([torch.tensor([1, 2, 3]), torch.tensor([4])])
    res = emb(nt)
    assert res.is_nested and [tuple(t.shape) for t in res.unbind()] == [(3, 4), (1, 4)]


def test_fit_with_model_checkpoint(cpu_context):
    """Keras-style loop: model.fit(data, optimizer, loss, epochs, callbacks=[ModelCheckpoint]) then reload"""
    import os
    import tempfile
    import openembedding_b200.torch as embed

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(200, 4)
            self.out = torch.nn.Linear(4, 1)

        def forward(self, ids):
            return self.out(self.emb(ids)).squeeze(-1)

    torch.manual_seed(0)
    model = embed.distributed_model(Net(), sparse_as_dense_size=0)
    opt = embed.distributed_optimizer(torch.optim.Adagrad(model.parameters(), lr=0.5))
    ids = torch.arange(200)
    y = (ids % 2).float()
    data = [(ids[i:i + 50], y[i:i + 50]) for i in range(0, 200, 50)]
    d = tempfile.mkdtemp()
    hist = model.fit(data, opt, torch.nn.functional.binary_cross_entropy_with_logits, epochs=6,
                     callbacks=[embed.ModelCheckpoint(d + "/ck{epoch}")])
    assert len(hist["loss"]) == 6 and hist["loss"][-1] < hist["loss"][0]
    assert os.path.exists(d + "/ck6") and os.path.exists(d + "/ck6.openembedding/openembedding/model_meta")
    want = model(ids).detach().clone()
    model.fit(data, opt, torch.nn.functional.binary_cross_entropy_with_logits, epochs=1)      # move away ...
    assert not torch.allclose(model(ids).detach(), want)
    model.load_weights(d + "/ck6")                                                             # ... and come back
    assert torch.allclose(model(ids).detach(), want, atol=1e-6)


@pytest.mark.parametrize("name,kw,tol", [
    ("SGD", dict(lr=0.1), 1e-6),
    ("SGD", dict(lr=0.1, momentum=0.9), 1e-5),
    ("SGD", dict(lr=0.1, momentum=0.9, nesterov=True), 1e-5),
    ("Adagrad", dict(lr=0.1, initial_accumulator_value=0.1, eps=1e-7), 1e-5),
    ("Adam", dict(lr=0.01, betas=(0.9, 0.999), eps=1e-7), 2e-3),
    ("Adamax", dict(lr=0.01, betas=(0.9, 0.999), eps=1e-7), 2e-3),
    ("RMSprop", dict(lr=0.01, alpha=0.9, eps=1e-7), 2e-3),
    ("Adadelta", dict(lr=1.0, rho=0.95, eps=1e-6), 2e-3),
])
def test_distributed_optimizer_matches_torch_dense(cpu_context, name, kw, tol):
    """hyper-parameter translation torch.optim.X -> server optimizer: a server Embedding trained through
    distributed_optimizer follows a dense nn.Embedding trained with the same torch optimizer (rows touched every step)"""
    import openembedding_b200.torch as embed
    torch.manual_seed(0)
    w0 = torch.randn(20, 5) * 0.1
    dense = torch.nn.Embedding(20, 5)
    dense.weight.data.copy_(w0)
    opt_d = getattr(torch.optim, name)(dense.parameters(), **kw)
    emb = embed.Embedding(20, 5, embeddings_initializer={"category": "constant", "value": 0.0})
    ids = torch.arange(20)
    emb.variable.variable  # created
    from openembedding_b200.context import get_context
    get_context().backend.load_rows(emb.variable.variable, ids.numpy().astype("uint64"), w0.numpy(),
                                    __import__("numpy").empty((20, 0), "float32"))
    opt_s = embed.distributed_optimizer(getattr(torch.optim, name)(emb.parameters(), **kw))
    target = torch.randn(20, 5)
    for _ in range(5):
        for m, o in ((dense, opt_d), (emb, opt_s)):
            loss = ((m(ids) - target) ** 2).sum()
            o.zero_grad(); loss.backward(); o.step()
    err = (emb(ids).detach() - dense.weight.detach()).abs().max().item()
    assert err < tol, (name, kw, err)
