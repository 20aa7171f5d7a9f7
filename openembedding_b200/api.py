"""User-facing API on ``torch.nn`` -- the counterpart of ``openembedding.tensorflow``
(openembedding/tensorflow/exb.py, 706 lines).

=====================  =================================================================
reference (exb.py)      here
=====================  =================================================================
Variable :222-360       ``Variable`` / ``distributed_variable``
Embedding :388-443      ``Embedding`` (an ``nn.Module``)
distributed_optimizer   ``distributed_optimizer`` + ``Adadelta..SGD`` classes (:446-488)
distributed_model       ``distributed_model`` / ``Model`` (:551-642)
save/load_server_model  same names (:491-503); ``save_as_original_model`` (:506-547)
pulling :645-691        ``pulling`` (input-pipeline prefetch)
persist/restore :697+   same names (host-tier lightweight checkpoint)
=====================  =================================================================

The reference makes TensorFlow call the PS by attaching a dummy ``[1, dim]`` variable to
each table and returning a fake gradient for it; the same trick is used here so that any
``torch.optim`` optimizer wrapped by ``distributed_optimizer`` drives the sparse update
in ``step()``.
"""
import copy
import math
import os
import shutil

import torch
from torch import nn

from . import checkpoint as _ckpt
from . import flags  # noqa: F401  (re-export like `from openembedding import *`)
from .config import HASH_KEY_RANGE, normalize_initializer, normalize_optimizer, str_dict
from .utils import timers
from .context import get_context

_HASH_KEY_RANGE = HASH_KEY_RANGE


# --------------------------------------------------------------------------- configs
def _torch_optimizer_config(optimizer, explicit=True):
    """torch.optim instance -> server optimizer config (keras names; exb.py:66-86)."""
    if isinstance(optimizer, dict):
        return normalize_optimizer(optimizer)
    if not isinstance(optimizer, torch.optim.Optimizer):
        raise ValueError("error optimizer: " + str(optimizer))
    g = optimizer.param_groups[0] if optimizer.param_groups else optimizer.defaults
    name = type(optimizer).__name__.lower()
    for base in type(optimizer).__mro__:
        if base.__module__.startswith("torch.optim") and base is not torch.optim.Optimizer:
            name = base.__name__.lower()
            break
    if getattr(optimizer, "_keras_category", None):
        name = optimizer._keras_category
    if name in ("adam", "adamw") and g.get("amsgrad", False):
        if explicit:
            raise ValueError("not support adam with amsgrad")
    if name == "rmsprop" and g.get("centered", False):
        if explicit:
            raise ValueError("not support centered rmsprop")
    if float(g.get("lr_decay", 0.0) or 0.0) != 0.0 and explicit:
        raise ValueError("not support learning rate decay")
    if float(g.get("weight_decay", 0.0) or 0.0) != 0.0 and explicit:
        raise ValueError("not support weight_decay on server side embeddings")
    lr = float(g.get("lr", 0.001))
    if name == "adagrad":
        return normalize_optimizer({"category": "adagrad", "learning_rate": lr,
                                    "initial_accumulator_value": g.get("initial_accumulator_value", 0.0),
                                    "epsilon": g.get("eps", 1e-10)})
    if name in ("adam", "adamw"):
        b1, b2 = g.get("betas", (0.9, 0.999))
        return normalize_optimizer({"category": "adam", "learning_rate": lr, "beta_1": b1, "beta_2": b2,
                                    "epsilon": g.get("eps", 1e-8)})
    if name == "adamax":
        b1, b2 = g.get("betas", (0.9, 0.999))
        return normalize_optimizer({"category": "adamax", "learning_rate": lr, "beta_1": b1, "beta_2": b2,
                                    "epsilon": g.get("eps", 1e-8)})
    if name == "adadelta":
        return normalize_optimizer({"category": "adadelta", "learning_rate": lr, "rho": g.get("rho", 0.9),
                                    "epsilon": g.get("eps", 1e-6)})
    if name == "rmsprop":
        return normalize_optimizer({"category": "rmsprop", "learning_rate": lr, "rho": g.get("alpha", 0.99),
                                    "momentum": g.get("momentum", 0.0), "epsilon": g.get("eps", 1e-8)})
    if name == "sgd":
        return normalize_optimizer({"category": "sgd", "learning_rate": lr, "momentum": g.get("momentum", 0.0),
                                    "nesterov": bool(g.get("nesterov", False))})
    if name == "ftrl":
        return normalize_optimizer({"category": "ftrl", "learning_rate": lr,
                                    "initial_accumulator_value": g.get("initial_accumulator_value", 0.1),
                                    "l1_regularization_strength": g.get("l1_regularization_strength", 0.0),
                                    "l2_regularization_strength": g.get("l2_regularization_strength", 0.0),
                                    "l2_shrinkage_regularization_strength": g.get("l2_shrinkage_regularization_strength", 0.0),
                                    "learning_rate_power": g.get("learning_rate_power", -0.5),
                                    "beta": g.get("beta", 0.0)})
    # Nadam etc.: wrapped by the reference too, but no server implementation exists
    # (EmbeddingOptimizer.h:393-395) -> fail like its factory does.
    raise ValueError("unsupported server optimizer: " + name)


# --------------------------------------------------------------------------- Variable
class _SparseRead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, graph_var, indices, variable):
        ctx.variable = variable
        ctx.indices = indices
        rows = variable._pull(indices)
        return rows

    @staticmethod
    def backward(ctx, grad):
        v = ctx.variable
        v._push(ctx.indices, grad)
        return torch.zeros_like(v.graph_var), None, None


class Variable:
    """A server-side (sharded, HBM-resident) embedding variable.

    ``shape[0] == -1`` selects the hashed 2**63 key space (exb.py:231-233).
    """

    def __init__(self, initializer=None, trainable=None, name=None, dtype=None, shape=None, num_shards=None,
                 sparse_as_dense=False, graph_var=None, host_tier_rows=None, host_store_rows=None):
        if not num_shards:
            num_shards = -1
        if shape is not None:
            shape = list(shape)
        if shape is not None and shape[0] == -1:
            shape[0] = _HASH_KEY_RANGE
            sparse_as_dense = False
        if graph_var is not None:
            if name or trainable is not None:
                raise ValueError("parameter conflict with graph_var")
            if shape and list(shape[1:]) != list(graph_var.shape[1:]):
                raise ValueError("graph_var shape not match")
        self._initialized = False
        self._name = name
        ctx = get_context()
        if dtype is None:
            dtype = torch.float32
        self._tdtype = dtype if isinstance(dtype, torch.dtype) else getattr(torch, str(dtype))
        if sparse_as_dense:
            if graph_var is None:
                init = normalize_initializer(initializer)
                w = torch.empty(shape, dtype=self._tdtype, device=ctx.device)
                _dense_init(w, init)
                graph_var = nn.Parameter(w, requires_grad=trainable is not False)
            self._sparse_as_dense = True
            self._shape = list(graph_var.shape)
            self.graph_var = graph_var
            return
        if shape is None or shape[0] <= 0 or shape[0] > _HASH_KEY_RANGE:
            raise ValueError("error shape")
        if not isinstance(initializer, dict):
            initializer = normalize_initializer(initializer)
        dtype_name = str(self._tdtype).replace("torch.", "")
        if graph_var is None:
            graph_var = nn.Parameter(torch.zeros([1] + shape[1:], dtype=self._tdtype, device=ctx.device),
                                     requires_grad=trainable is not False)
        embedding_dim = 1
        for d in shape[1:]:
            embedding_dim *= d
        self._sparse_as_dense = False
        self._shape = shape
        self._initialized = True
        self.graph_var = graph_var
        self.storage = ctx.create_storage(num_shards)
        tiered = host_tier_rows is not None or bool(ctx.env["server"]["host_tier_root_path"]) \
            or bool(ctx.env["server"]["pmem_pool_root_path"])
        self.variable = ctx.create_variable(self.storage, shape[0], embedding_dim, dtype_name, force_hash=tiered,
                                            capacity=(2 * host_tier_rows) if host_tier_rows else None)
        ctx.set_initializer(self.variable, initializer)
        self.tier = None
        if tiered:
            from .host_tier import make_tiered
            self.tier = make_tiered(ctx, self.variable, host_tier_rows, host_rows=host_store_rows)
        self.model_uuid = ctx.model_uuid
        self.optimizer_set = False
        self._prefetched = []
        ctx.tracks[id(graph_var)] = self

    # -- properties
    @property
    def name(self):
        return self._name or ("variable_%d" % self.variable.variable_id if not self._sparse_as_dense else "dense")

    @property
    def shape(self):
        return self._shape

    @property
    def sparse_as_dense(self):
        return self._sparse_as_dense

    # -- data path
    def _pull(self, indices):
        ctx = get_context()
        flat = indices.reshape(-1)
        if self.tier is not None:
            self.tier.prefetch(flat)          # promote missing rows from the host store into HBM
        with timers.vtimer(1, "client", "pull_weights", cuda=ctx.device.type == "cuda"):
            rows = ctx.backend.pull(self.variable, flat)
        if timers.enabled:
            timers.add_count("pull_indices", flat.numel())
            if not flat.is_cuda:      # on the GPU the engine counts unique rows itself (engine.status(): pull_unique /
                timers.add_count("pull_unique", int(torch.unique(flat).numel()))   # update_unique): no sync here
        return rows.reshape(tuple(indices.shape) + tuple(self._shape[1:])).to(self._tdtype)

    def _push(self, indices, grads):
        ctx = get_context()
        with timers.vtimer(1, "client", "push_gradients", cuda=ctx.device.type == "cuda"):
            ctx.backend.push(self.variable, indices.reshape(-1), grads.reshape(-1, self.variable.dim))
        if self.tier is not None:
            self.tier.mark_updated(indices)

    def prefetch(self, indices, steps=None):
        """Stage the ids of a future batch on the device ahead of time.

        The reference parks a future-batch pull on the server until ``batch_id`` catches up
        (exb_ops.cpp:139-175, EmbeddingPullOperator.cpp:117-145). With tables in HBM the row
        gather itself is cheap; what is worth overlapping is the host->device copy of ids."""
        if self.sparse_as_dense:
            raise ValueError("should not prefetch for sparse as dense.")
        ctx = get_context()
        if ctx.device.type == "cuda" and not indices.is_cuda:
            src = indices if indices.is_pinned() else indices.pin_memory()
            return src.to(ctx.device, non_blocking=True)
        return indices

    def sparse_read(self, indices):
        if self.sparse_as_dense:
            return nn.functional.embedding(indices.to(self.graph_var.device), self.graph_var)
        ctx = get_context()
        if indices.dtype != torch.int64:
            indices = indices.to(torch.int64)
        indices = indices.to(ctx.device)
        if torch.is_grad_enabled() and self.graph_var.requires_grad:
            return _SparseRead.apply(self.graph_var, indices, self)
        return self._pull(indices)

    pull_weights = sparse_read

    def set_server_optimizer(self, optimizer):
        if self.sparse_as_dense:
            raise ValueError("no need optimizer for sparse as dense.")
        if not isinstance(optimizer, dict):
            optimizer = _torch_optimizer_config(optimizer)
        get_context().set_optimizer(self.variable, optimizer)
        self.optimizer_set = True

    def push_gradients(self, indices, gradients):
        if self.sparse_as_dense:
            raise ValueError("no need update weights for sparse as dense.")
        self._push(indices.to(torch.int64), gradients)
        return torch.zeros_like(self.graph_var)

    def update_weights(self, fake_grad=None):
        if self.sparse_as_dense:
            raise ValueError("no need update weights for sparse as dense.")
        ctx = get_context()
        with timers.vtimer(1, "client", "update_weights", cuda=ctx.device.type == "cuda"):
            ctx.backend.update([self.variable])
        if self.tier is not None:
            self.tier.next_work()

    def _finalize(self):
        self._initialized = False


distributed_variable = Variable


def _dense_init(w, init):
    c = init["category"]
    with torch.no_grad():
        if c == "constant":
            w.fill_(init["value"])
        elif c == "uniform":
            w.uniform_(init["minval"], init["maxval"])
        else:
            w.normal_(init["mean"], init["stddev"])


# --------------------------------------------------------------------------- Embedding
class Embedding(nn.Module):
    """Drop-in for ``nn.Embedding`` / ``keras.layers.Embedding`` backed by the sharded engine.

    input_dim: vocabulary size; ``-1``/``None`` -> ids in [0, 2**63) (hash table).
    sparse_as_dense: keep the table as an ordinary replicated parameter with dense
    gradients (the "cache" mode of the reference, exb.py:241-248).
    """

    def __init__(self, input_dim, output_dim, embeddings_initializer="uniform", embeddings_regularizer=None,
                 activity_regularizer=None, embeddings_constraint=None, mask_zero=False, input_length=None,
                 num_shards=None, sparse_as_dense=False, explicit=True, dtype=None, name=None, host_tier_rows=None,
                 host_store_rows=None):
        super().__init__()
        if input_dim is None:
            input_dim = -1
        if input_dim == -1:
            input_dim = _HASH_KEY_RANGE
            sparse_as_dense = False
        if not sparse_as_dense and explicit:
            if embeddings_regularizer:
                raise ValueError("not support embeddings_regularizer")
            if embeddings_constraint:
                raise ValueError("not support embeddings_constraint")
        if not activity_regularizer:
            activity_regularizer = embeddings_regularizer
        if input_dim <= 0:
            raise ValueError("error input_dim")
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.num_shards = num_shards
        self.sparse_as_dense = bool(sparse_as_dense)
        self.activity_regularizer = activity_regularizer
        self.mask_zero, self.input_length = mask_zero, input_length
        self.layer_name = name
        self.embeddings_initializer = embeddings_initializer
        self.server_initializer = normalize_initializer(embeddings_initializer, explicit=explicit)
        ctx = get_context()
        dtype = dtype or torch.float32
        if self.sparse_as_dense:
            w = torch.empty((self.input_dim, self.output_dim), dtype=dtype, device=ctx.device)
            _dense_init(w, self.server_initializer)
            self.embeddings = nn.Parameter(w)
            self.variable = Variable(sparse_as_dense=True, graph_var=self.embeddings)
        else:
            self.embeddings = nn.Parameter(torch.zeros((1, self.output_dim), dtype=dtype, device=ctx.device))
            self.variable = Variable(initializer=dict(self.server_initializer), dtype=dtype,
                                     shape=(self.input_dim, self.output_dim), num_shards=num_shards,
                                     graph_var=self.embeddings, host_tier_rows=host_tier_rows,
                                     host_store_rows=host_store_rows)
        self.built = True

    def forward(self, inputs, offsets=None, per_sample_weights=None, mode="sum"):
        """``inputs``: int tensor of any shape -> ``inputs.shape + (output_dim,)``.

        Multi-hot / ragged features (the reference reaches them through
        ``tf.ragged.map_flat_values(embedding, ragged_ids)``, exb.py:388-443): pass the flat
        values as a 1-D ``inputs`` plus ``offsets`` (start of every bag, like
        ``nn.EmbeddingBag``) and get one pooled row per bag (``mode`` sum | mean | max is
        applied on this rank after ONE pull of the flat ids, so duplicated ids inside a batch
        travel once). A ``torch.nested`` tensor is accepted too and returns a nested result."""
        grp = getattr(self, "_group", None)
        if grp is not None and offsets is None and not getattr(inputs, "is_nested", False):
            hit = grp.lookup(self, inputs)
            if hit is not None:
                if self.activity_regularizer is not None and self.training:
                    self.activity_loss = self.activity_regularizer(hit)
                return hit
        if offsets is None and getattr(inputs, "is_nested", False):
            parts = inputs.unbind()
            flat = torch.cat([p.reshape(-1) for p in parts])
            rows = self.variable.sparse_read(flat)
            out, o = [], 0
            for p in parts:
                out.append(rows[o:o + p.numel()].reshape(tuple(p.shape) + (self.output_dim,)))
                o += p.numel()
            return torch.nested.nested_tensor(out)
        out = self.variable.sparse_read(inputs)
        if offsets is not None:
            out = _pool_bags(out, offsets.to(out.device), per_sample_weights, mode)
        if self.activity_regularizer is not None and self.training:
            self.activity_loss = self.activity_regularizer(out)
        return out

    def extra_repr(self):
        v = "2**63" if self.input_dim >= _HASH_KEY_RANGE else str(self.input_dim)
        return "%s, %d, sparse_as_dense=%s" % (v, self.output_dim, self.sparse_as_dense)


def _pool_bags(rows, offsets, per_sample_weights=None, mode="sum"):
    """rows [n, D] of the flat ids, offsets [bags] -> [bags, D] (differentiable)."""
    n, bags = rows.shape[0], offsets.numel()
    if per_sample_weights is not None:
        if mode != "sum":
            raise ValueError("per_sample_weights needs mode='sum'")
        rows = rows * per_sample_weights.to(rows.device, rows.dtype).reshape(-1, 1)
    ends = torch.cat([offsets[1:], torch.tensor([n], device=offsets.device, dtype=offsets.dtype)])
    lens = (ends - offsets).clamp_(min=0)
    bag_of = torch.repeat_interleave(torch.arange(bags, device=rows.device), lens.to(rows.device))
    if mode in ("sum", "mean"):
        out = torch.zeros((bags, rows.shape[1]), dtype=rows.dtype, device=rows.device).index_add_(0, bag_of, rows)
        if mode == "mean":
            out = out / lens.clamp(min=1).to(rows.dtype).reshape(-1, 1)
        return out
    if mode == "max":
        out = torch.full((bags, rows.shape[1]), float("-inf"), dtype=rows.dtype, device=rows.device)
        out = out.scatter_reduce(0, bag_of.reshape(-1, 1).expand_as(rows), rows, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    raise ValueError("mode must be sum, mean or max")


# --------------------------------------------------------------------------- optimizers
class Ftrl(torch.optim.Optimizer):
    """Dense-parameter FTRL with tf.keras semantics (the formulas of
    openembedding/variable/EmbeddingOptimizer.h:230-293, shared with the server kernels)."""
    _keras_category = "ftrl"

    def __init__(self, params, lr=0.001, initial_accumulator_value=0.1, l1_regularization_strength=0.0,
                 l2_regularization_strength=0.0, l2_shrinkage_regularization_strength=0.0,
                 learning_rate_power=-0.5, beta=0.0):
        defaults = dict(lr=lr, initial_accumulator_value=initial_accumulator_value,
                        l1_regularization_strength=l1_regularization_strength,
                        l2_regularization_strength=l2_regularization_strength,
                        l2_shrinkage_regularization_strength=l2_shrinkage_regularization_strength,
                        learning_rate_power=learning_rate_power, beta=beta)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for g in self.param_groups:
            lr, l1, l2 = g["lr"], g["l1_regularization_strength"], g["l2_regularization_strength"]
            l2s, p, beta = g["l2_shrinkage_regularization_strength"], -g["learning_rate_power"], g["beta"]
            adj_l2 = l2 + beta / lr / 2
            for w in g["params"]:
                if w.grad is None:
                    continue
                st = self.state[w]
                if not st:
                    st["accum"] = torch.full_like(w, g["initial_accumulator_value"])
                    st["linear"] = torch.zeros_like(w)
                grad = w.grad
                gg = grad + 2 * l2s * w
                accum_new = st["accum"] + grad * grad
                sigma = (accum_new.pow(p) - st["accum"].pow(p)) / lr
                st["linear"] += gg - sigma * w
                st["accum"] = accum_new
                quadratic = accum_new.pow(p) / lr + 2 * adj_l2
                l1_adj = st["linear"].clamp(-l1, l1)
                w.copy_((l1_adj - st["linear"]) / quadratic)
        return loss


_DistributedOptimizerClass = {}


def _DistributedOptimizer(T):
    """Subclass of optimizer class T whose ``step`` also drives the server-side update
    (reference: exb.py:446-468 overriding ``_resource_apply_dense``)."""
    if T in _DistributedOptimizerClass:
        return _DistributedOptimizerClass[T]

    class _Optimizer(T):
        def __init__(self, *args, explicit=True, **kwargs):
            super().__init__(*args, **kwargs)
            self._explicit = explicit
            self.server_optimizer = _torch_optimizer_config(self, explicit=explicit)

        @torch.no_grad()
        def step(self, closure=None):
            ctx = get_context()
            tracked, stash = [], []
            for g in self.param_groups:
                for p in g["params"]:
                    v = ctx.tracks.get(id(p))
                    if v is not None and not v.sparse_as_dense:
                        if not v.optimizer_set:
                            v.set_server_optimizer(self.server_optimizer)
                        if p.grad is not None:
                            tracked.append(v)
                        stash.append((p, p.grad))
                        p.grad = None          # the dummy [1, dim] parameter is never updated locally
            for grp in getattr(ctx, "groups", []):
                grp.flush()                                # ONE push+update launch for every grouped Embedding
            if tracked:
                ctx.backend.update([v.variable for v in tracked])
                for v in tracked:
                    if v.tier is not None:
                        v.tier.next_work()
            ctx.step_done()
            try:
                return super().step(closure)
            finally:
                for p, gr in stash:
                    p.grad = gr

    _Optimizer.__name__ = T.__name__
    _Optimizer.__qualname__ = T.__qualname__
    _DistributedOptimizerClass[T] = _Optimizer
    return _Optimizer


Adadelta = _DistributedOptimizer(torch.optim.Adadelta)
Adagrad = _DistributedOptimizer(torch.optim.Adagrad)
Adam = _DistributedOptimizer(torch.optim.Adam)
Adamax = _DistributedOptimizer(torch.optim.Adamax)
Nadam = _DistributedOptimizer(torch.optim.NAdam)   # constructing it raises: no server implementation
RMSprop = _DistributedOptimizer(torch.optim.RMSprop)
SGD = _DistributedOptimizer(torch.optim.SGD)
FtrlDistributed = _DistributedOptimizer(Ftrl)


def distributed_optimizer(optimizer, explicit=True):
    """Return a distributed optimizer that trains dense parameters locally and the
    server-side embeddings through the sparse engine. If a data-parallel wrapper is
    used (DDP / ``parallel.allreduce``), create it after this call (exb.py:481-488)."""
    cls = _DistributedOptimizer(type(optimizer))
    new = cls.__new__(cls)
    new.__dict__.update(optimizer.__dict__)
    new._explicit = explicit
    new.server_optimizer = _torch_optimizer_config(optimizer, explicit=explicit)
    return new


# --------------------------------------------------------------------------- model save / load
def save_server_model(model, filepath, include_optimizer=True):
    """Save the parameters held by the sparse engine. Collective: every rank writes its
    own shard (there are no server processes to do it on a single caller's behalf)."""
    ctx = get_context()
    _ckpt.save_model(ctx, filepath, include_optimizer=include_optimizer)


def load_server_model(model, filepath):
    """Load the server parameters. Must be called synchronously by all workers."""
    _ckpt.load_model(get_context(), filepath)


def _iter_embeddings(model):
    for name, mod in model.named_modules():
        if isinstance(mod, Embedding):
            yield name, mod


def _to_original(model):
    """Deep copy of `model` where every server Embedding became a plain nn.Embedding."""
    for _, layer in _iter_embeddings(model):
        if layer.variable.shape[0] >= _HASH_KEY_RANGE:
            raise ValueError("can not convert sparse variable to nn.Embedding.")
    ctx = get_context()
    memo = {}
    for _, layer in _iter_embeddings(model):   # do not deep-copy engine handles
        memo[id(layer)] = layer
    grp = getattr(model, "_embedding_group", None)
    if grp is not None:
        memo[id(grp)] = grp
    clone = copy.deepcopy(model, memo)
    if grp is not None:                        # the export is a plain model: no fused-group hooks, no engine state
        for hooks in (clone._forward_pre_hooks, clone._forward_hooks):
            for k in [k for k, h in hooks.items() if getattr(h, "__self__", None) is grp]:
                del hooks[k]
        for d in (getattr(clone, "_forward_pre_hooks_with_kwargs", None), getattr(clone, "_forward_hooks_with_kwargs", None)):
            if isinstance(d, dict):
                for k in [k for k in d if k not in clone._forward_pre_hooks and k not in clone._forward_hooks]:
                    del d[k]
        if "_embedding_group" in clone.__dict__:
            del clone.__dict__["_embedding_group"]

    def convert(parent):
        for cname, child in list(parent.named_children()):
            if isinstance(child, Embedding):
                plain = nn.Embedding(child.input_dim, child.output_dim)
                if child.sparse_as_dense:
                    plain.weight.data.copy_(child.embeddings.data)
                else:
                    batch = 2 ** 20 // child.output_dim + 1      # exb.py:541
                    with torch.no_grad():
                        for i in range(0, child.input_dim, batch):
                            idx = torch.arange(i, min(child.input_dim, i + batch), device=ctx.device)
                            plain.weight.data[i:i + idx.numel()] = child.variable.sparse_read(idx).to("cpu")
                setattr(parent, cname, plain)
            else:
                convert(child)

    convert(clone)
    if isinstance(clone, Embedding):
        raise ValueError("wrap the Embedding in a module to export it")
    return clone


def save_as_original_model(model, filepath, overwrite=True, include_optimizer=False, **kwargs):
    """Export a stand-alone PyTorch model (plain ``nn.Embedding`` weights) that needs no
    engine to serve -- counterpart of exb.py:506-547."""
    if include_optimizer is True:
        raise ValueError("not support include optimizer")
    ctx = get_context()
    if os.path.exists(filepath) and not overwrite:
        raise IOError("exists: " + filepath)
    # COLLECTIVE (unlike the reference, whose server processes answer a single caller): the export pulls every
    # row through the sharded engine, which every rank has to take part in. All ranks run the pulls, only
    # rank 0 writes the file, and a barrier makes the file visible before anybody returns.
    clone = _to_original(model)
    if type(clone).__name__ in _DistributedModelNames:
        clone.__class__ = clone._Class_base
    if ctx.rank == 0:
        d = os.path.dirname(os.path.abspath(filepath))
        os.makedirs(d, exist_ok=True)
        torch.save(clone.cpu(), filepath)
    ctx.barrier()
    return clone


_DistributedModelClass = {}
_DistributedModelNames = set()


def _DistributedModel(T):
    if T in _DistributedModelClass:
        return _DistributedModelClass[T]

    class _Model(T):
        _Class_base = T

        def save(self, filepath, overwrite=True, include_optimizer=True, **kwargs):
            """dense state_dict -> <filepath>/model.pt ; server tables -> <filepath>/openembedding/"""
            ctx = get_context()
            if ctx.rank == 0:
                os.makedirs(filepath, exist_ok=True)
                torch.save(self.state_dict(), os.path.join(filepath, "model.pt"))
                if os.path.exists(filepath + "/openembedding"):
                    shutil.rmtree(filepath + "/openembedding")
            ctx.barrier()
            save_server_model(self, filepath + "/openembedding", include_optimizer=include_optimizer)

        def save_weights(self, filepath, **kwargs):
            ctx = get_context()
            if ctx.rank == 0:
                d = os.path.dirname(os.path.abspath(filepath))
                os.makedirs(d, exist_ok=True)
                torch.save(self.state_dict(), filepath)
                if os.path.exists(filepath + ".openembedding/openembedding"):
                    shutil.rmtree(filepath + ".openembedding/openembedding")
            ctx.barrier()
            save_server_model(self, filepath + ".openembedding/openembedding", include_optimizer=True)

        def load_weights(self, filepath, **kwargs):
            if os.path.isdir(filepath) and os.path.exists(os.path.join(filepath, "model.pt")):
                self.load_state_dict(torch.load(os.path.join(filepath, "model.pt"), map_location="cpu"))
                load_server_model(self, filepath + "/openembedding")
                return
            self.load_state_dict(torch.load(filepath, map_location="cpu"))
            if os.path.exists(filepath + ".openembedding/openembedding"):
                load_server_model(self, filepath + ".openembedding/openembedding")
            elif os.path.exists(os.path.split(filepath)[0] + "/../openembedding"):
                load_server_model(self, os.path.split(filepath)[0] + "/../openembedding")
            else:
                raise IOError("embed not exist: " + filepath)

        def save_as_original_model(self, filepath, *args, **kwargs):
            return save_as_original_model(self, filepath, *args, **kwargs)

        def fit(self, data, optimizer, loss_fn, epochs=1, callbacks=(), steps_per_epoch=None, verbose=0):
            """Minimal Keras-style training loop so that scripts written against the reference's
            ``model.fit(dataset, epochs=, callbacks=[ModelCheckpoint(...)])`` port one to one
            (examples/criteo_deepctr_network.py:60-70 of the reference). ``data`` yields
            ``(inputs, labels)``; ``inputs`` may be a tensor, a tuple/list (positional) or a dict
            (keyword arguments of ``forward``). Returns ``{"loss": [per-epoch mean]}``."""
            history = {"loss": []}
            for cb in callbacks:
                cb.set_model(self)
            for epoch in range(epochs):
                tot, n = 0.0, 0
                for step, (inputs, labels) in enumerate(data):
                    if steps_per_epoch is not None and step >= steps_per_epoch:
                        break
                    if isinstance(inputs, dict):
                        out = self(**inputs)
                    elif isinstance(inputs, (tuple, list)):
                        out = self(*inputs)
                    else:
                        out = self(inputs)
                    loss = loss_fn(out, labels.to(out.device))
                    optimizer.zero_grad()
                    loss.backward()
                    optimizer.step()
                    tot += float(loss.detach())
                    n += 1
                history["loss"].append(tot / max(n, 1))
                if verbose and get_context().rank == 0:
                    print("Epoch %d/%d - loss: %.4f" % (epoch + 1, epochs, history["loss"][-1]))
                for cb in callbacks:
                    cb.on_epoch_end(epoch, {"loss": history["loss"][-1]})
            return history

    _Model.__name__ = T.__name__
    _Model.__qualname__ = T.__qualname__
    _DistributedModelClass[T] = _Model
    _DistributedModelNames.add(T.__name__)
    return _Model


Model = _DistributedModel(nn.Module)


class ModelCheckpoint:
    """``ModelCheckpoint(filepath, save_weights_only=True)`` for ``Model.fit``: after every epoch calls
    ``model.save_weights`` (server tables INCLUDING optimizer state, like the reference's Keras callback
    path, exb.py:563-567) or ``model.save``; ``{epoch}`` in ``filepath`` is replaced by the 1-based epoch."""

    def __init__(self, filepath, save_weights_only=True, include_optimizer=True):
        self.filepath, self.save_weights_only, self.include_optimizer = filepath, save_weights_only, include_optimizer
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        path = self.filepath.format(epoch=epoch + 1)
        if self.save_weights_only:
            self.model.save_weights(path)
        else:
            self.model.save(path, include_optimizer=self.include_optimizer)


class _GroupPull(torch.autograd.Function):
    """ONE pull launch for every server Embedding of a model; the gradient rows are kept for ONE fused
    push+update at ``optimizer.step()`` (reference: 26 PullWeights + 26 PushGradients ops per step,
    exb_ops.cpp:208-414)."""

    @staticmethod
    def forward(ctx, anchor, ids, group):
        ctx.group = group
        ctx.ids = ids
        return group.plan.pull(ids, train=torch.is_grad_enabled())

    @staticmethod
    def backward(ctx, grad):
        g = ctx.group
        g.pending = (ctx.ids, grad.contiguous())
        return torch.zeros_like(g.anchor), None, None


class EmbeddingGroup:
    """Fused sparse path behind the public API (``distributed_model``).

    All server-side ``Embedding`` layers of a model whose indices are fed DIRECTLY by a model input -- a whole
    1-D / ``[B, 1]`` input tensor (positional, keyword or dict entry) or one column of a 2-D integer input, the
    same condition as the reference's ``pulling`` (exb.py:645-691: "Embedding fed directly by exactly one
    InputLayer") -- are served by ONE ``SparsePlan``: a forward pre-hook pulls all of them in one launch, every
    layer's ``forward`` returns its slice, and ``distributed_optimizer.step()`` applies all their gradients in one
    push+update launch. The mapping input -> layer is discovered by tracing the first forward pass; layers that do
    not qualify (derived indices, multi-hot bags, host-tier tables, ``sparse_as_dense``) keep the per-variable path."""

    def __init__(self, model):
        self.model = model
        self.ctx = get_context()
        self.layers = [l for _, l in _iter_embeddings(model)
                       if not l.sparse_as_dense and getattr(l.variable, "tier", None) is None
                       and l.variable._tdtype == torch.float32]
        for l in self.layers:
            l._group = self
        self.state = "trace" if (self.layers and self.ctx.device.type == "cuda") else "off"
        self.trace = []             # (layer, indices tensor) seen during the traced forward
        self.members = []           # [(layer, source)] source = (kind, key, column)
        self.plan = None
        self.cache = {}
        self.pending = None
        self.anchor = None
        self.batch = None
        self.ctx.groups = getattr(self.ctx, "groups", [])
        self.ctx.groups.append(self)
        model.register_forward_pre_hook(self._pre, with_kwargs=True)
        model.register_forward_hook(self._post, with_kwargs=True)

    # ---- inputs
    @staticmethod
    def _flatten_inputs(args, kwargs):
        out = []
        for i, a in enumerate(args):
            if isinstance(a, dict):
                out += [(("argdict", i, k), v) for k, v in a.items() if torch.is_tensor(v)]
            elif isinstance(a, (list, tuple)):
                out += [(("argseq", i, j), v) for j, v in enumerate(a) if torch.is_tensor(v)]
            elif torch.is_tensor(a):
                out.append((("arg", i, None), a))
        for k, v in kwargs.items():
            if torch.is_tensor(v):
                out.append((("kw", k, None), v))
        return out

    @staticmethod
    def _match(ind, inputs):
        """which model input (and column) is this indices tensor?"""
        if ind.dtype not in (torch.int64, torch.int32) or ind.dim() not in (1, 2):
            return None
        if ind.dim() == 2 and ind.shape[1] != 1:
            return None
        n = ind.shape[0]
        for key, t in inputs:
            if t.dtype != ind.dtype or t.device != ind.device or t.shape[0] != n:
                continue
            if t.dim() == ind.dim() and t.shape == ind.shape and t.data_ptr() == ind.data_ptr() and t.stride() == ind.stride():
                return key + (None,)
            if t.dim() == 2 and t.untyped_storage().data_ptr() == ind.untyped_storage().data_ptr():
                off = (ind.data_ptr() - t.data_ptr()) // t.element_size()
                if 0 <= off < t.shape[1] and ind.stride(0) == t.stride(0) and t.stride(1) == 1:
                    return key + (int(off),)
        return None

    def _fetch(self, args, kwargs, src):
        kind, a, b, col = src
        if kind == "arg":
            t = args[a]
        elif kind == "argdict":
            t = args[a][b]
        elif kind == "argseq":
            t = args[a][b]
        else:
            t = kwargs[a]
        return t if col is None else t[:, col]

    # ---- hooks
    def _pre(self, module, args, kwargs):
        self.cache = {}
        if self.state == "trace":
            self.trace = []
            self._inputs = self._flatten_inputs(args, kwargs)
            return None
        if self.state != "on" or not torch.is_grad_enabled() or not module.training:
            return None
        try:
            cols = [self._fetch(args, kwargs, src).reshape(-1) for _, src in self.members]
        except Exception:
            return None
        n = cols[0].shape[0]
        if n != self.batch or any(c.shape[0] != n for c in cols):
            return None                                   # odd batch (last one of an epoch): per-variable path
        ids = torch.stack([c.to(device=self.ctx.device, dtype=torch.int64) for c in cols], dim=1).contiguous()
        out = _GroupPull.apply(self.anchor, ids, self)
        for (layer, _), sl in zip(self.members, self.plan.feature_slices()):
            self.cache[id(layer)] = out[:, sl]
        return None

    def lookup(self, layer, inputs):
        if self.state == "trace":
            self.trace.append((layer, inputs))
            return None
        hit = self.cache.pop(id(layer), None)
        if hit is None:
            return None
        return hit.reshape(tuple(inputs.shape) + (layer.output_dim,)).to(layer.variable._tdtype)

    def _post(self, module, args, kwargs, output):
        if self.state != "trace":
            return None
        members, seen = [], set()
        for layer, ind in self.trace:
            src = self._match(ind, self._inputs) if torch.is_tensor(ind) else None
            if src is not None and id(layer) not in seen and layer.variable.graph_var.requires_grad:
                members.append((layer, src))
                seen.add(id(layer))
        self.trace, self._inputs = [], None
        if len(members) < 2:
            self.state = "off"
            return None
        n = self._fetch(args, kwargs, members[0][1]).reshape(-1).shape[0]
        be = self.ctx.backend
        metas = [l.variable.variable for l, _ in members]
        be.ensure_allocated(metas)
        self.plan = be.engine.make_plan([m.handle for m in metas], n)
        be.engine.connect(be.group)
        self.members, self.batch = members, n
        self.anchor = torch.zeros(1, device=self.ctx.device, requires_grad=True)
        self.state = "on"
        return None

    # ---- called by the distributed optimizer
    def flush(self):
        if self.pending is None:
            return False
        ids, grad = self.pending
        self.pending = None
        self.plan.push_update(ids, grad)
        return True


def distributed_model(model, sparse_as_dense_size=64, num_shards=None, override_method=True, explicit=False):
    """Replace every ``nn.Embedding`` of `model` by a server-side ``Embedding``
    (``sparse_as_dense`` when ``num_embeddings <= sparse_as_dense_size``) and add
    ``save / save_weights / load_weights / save_as_original_model`` (exb.py:593-642).
    Do not keep using the input model's old embedding modules afterwards."""
    def convert(parent):
        for cname, child in list(parent.named_children()):
            if isinstance(child, nn.Embedding) and not isinstance(child, Embedding):
                sad = child.num_embeddings <= sparse_as_dense_size
                new = Embedding(child.num_embeddings, child.embedding_dim,
                                embeddings_initializer={"category": "normal", "mean": 0.0, "stddev": 1.0},
                                num_shards=num_shards, sparse_as_dense=sad, explicit=explicit,
                                dtype=child.weight.dtype, name=cname)
                setattr(parent, cname, new)
            else:
                convert(child)

    if isinstance(model, nn.Embedding) and not isinstance(model, Embedding):
        raise ValueError("wrap the nn.Embedding in a module")
    convert(model)
    ctx = get_context()
    model.to(ctx.device)
    if override_method:
        model.__class__ = _DistributedModel(model.__class__)
    if os.environ.get("EXB_API_FUSED", "1") != "0":
        model._embedding_group = EmbeddingGroup(model)      # one pull / one push+update launch per step
    return model


# --------------------------------------------------------------------------- input pipeline
def pulling(dataset, model, steps=None):
    """EXPERIMENTAL: wrap an iterable of ``(inputs_dict, labels)`` batches so that the ids of
    every server Embedding that is fed directly by exactly one input column are staged
    on the device one batch ahead (exb.py:645-691). ``model.input_names`` (list of column
    names aligned with the Embedding modules via ``layer_name``) selects the columns."""
    emb_by_col = {}
    for _, layer in _iter_embeddings(model):
        if not layer.sparse_as_dense and layer.layer_name:
            emb_by_col.setdefault(layer.layer_name, []).append(layer)
    cols = {c: ls[0] for c, ls in emb_by_col.items() if len(ls) == 1}

    def gen():
        n = 0
        for batch in dataset:
            if steps is not None and n >= steps:
                return
            n += 1
            if isinstance(batch, (tuple, list)) and isinstance(batch[0], dict):
                feats = dict(batch[0])
                for c, layer in cols.items():
                    if c in feats:
                        feats[c] = layer.variable.prefetch(feats[c], steps=steps)
                yield (feats,) + tuple(batch[1:])
            else:
                yield batch

    class _Pulling:
        def __iter__(self):
            # one-batch lookahead so the H2D copy of batch k+1 overlaps step k
            it = gen()
            try:
                nxt = next(it)
            except StopIteration:
                return
            for cur in it:
                yield nxt
                nxt = cur
            yield nxt

    return _Pulling()


# --------------------------------------------------------------------------- host-tier persist (pmem experimental in the reference)
def should_persist_server_model(model):
    from .host_tier import should_persist
    return should_persist(get_context())


def persist_server_model(model, filepath, persist_pending_window):
    from .host_tier import persist_model
    persist_model(get_context(), filepath, persist_pending_window)


def restore_server_model(model, filepath):
    from .host_tier import restore_model
    restore_model(get_context(), filepath)
