"""Same-box anchor for the headline numbers: the SAME model / config / data as ``bench.py``, written the way
one would without this repo -- plain PyTorch:

* embedding tables row-sharded over the ranks (``id % W``), ids / rows / gradients exchanged with
  ``torch.distributed.all_to_all_single`` over NCCL (the "allreduce + all-to-all" recipe every PyTorch
  recommender uses; the reference's own comparison arm is Horovod allreduce-only, benchmark.md:5-35),
  per-step ``torch.unique`` de-duplication, sparse Adagrad with ``index_add_`` / indexed updates;
* dense model = ``torch.nn`` (cuBLAS GEMMs) under ``torch.autocast(bf16)`` or in fp32, FM in fp32;
* dense gradients: one flat buffer, ``dist.all_reduce`` (NCCL, sum like ``hvd.DistributedOptimizer(op=Sum)``),
  tf.keras-style Adagrad on the flat buffer.

None of this repo's kernels, engine or models are on this path (only ``bench.py``'s data generator and the
vocabulary lists are shared). It is NOT the reference (``--impl reference`` stays "unavailable": the reference
needs TensorFlow + Horovod + its CMake-built C++ stack); it is the strongest stock-library implementation of the
reference's benchmark we could write, so "x times the baseline" has a denominator measured on the same box.

The architecture mirrors ``models/fused_dense.FusedCTR`` (DeepCTR DeepFM / WDL: per sparse feature an
embedding of dim D and a linear weight, 13 dense features, DNN 400x3 / 512-256-128-32) and can be initialised
from one (``load_from_fused``) for the bf16-vs-fp32 loss-parity run (tools/loss_parity.py).
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn


class NcclBaselineCTR:
    def __init__(self, vocab, num_dense=13, embedding_dim=64, model="deepfm", batch=4096, cache_threshold=0,
                 compute_dtype=torch.bfloat16, lr=0.001, sparse_lr=0.001, init_acc=0.1, eps=1e-7, hidden=None,
                 rank=0, world=1, device=None, seed=0):
        self.vocab, self.nd, self.D, self.B = list(vocab), num_dense, embedding_dim, batch
        self.model = model.lower()
        assert self.model in ("deepfm", "wdl")
        self.rank, self.W = rank, world
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.compute_dtype = compute_dtype
        self.lr, self.slr, self.eps = lr, sparse_lr, eps
        nf = len(self.vocab)
        self.cached = [f for f, v in enumerate(self.vocab) if 0 < v < cache_threshold]
        self.server = [f for f in range(nf) if f not in self.cached]
        self.ns, self.nc = len(self.server), len(self.cached)
        # ---- server tables: every feature's local shard concatenated into ONE [rows, D+1] slab (emb | linear)
        W = world
        self.rows_f = [(self.vocab[f] + W - 1) // W for f in self.server]
        offs, o = [], 0
        for r in self.rows_f:
            offs.append(o)
            o += r
        self.R = max(o, 1)
        self.row_off = torch.tensor(offs or [0], dtype=torch.int64, device=self.dev)
        self.srv_cols = torch.tensor(self.server or [0], dtype=torch.int64, device=self.dev)
        self.table = torch.zeros((self.R, embedding_dim + 1), dtype=torch.float32, device=self.dev)
        self.table_acc = torch.full((self.R, embedding_dim + 1), init_acc, dtype=torch.float32, device=self.dev)
        # ---- dense part
        hidden = list(hidden or ((400, 400, 400) if self.model == "deepfm" else (512, 256, 128, 32)))
        self.hidden = hidden
        g = torch.Generator(device="cpu").manual_seed(seed)
        layers, prev = [], nf * embedding_dim + num_dense
        for h in hidden:
            lin = nn.Linear(prev, h)
            with torch.no_grad():
                lin.weight.copy_(torch.randn(h, prev, generator=g) * math.sqrt(2.0 / (prev + h)))
                lin.bias.zero_()
            layers += [lin, nn.ReLU()]
            prev = h
        self.dnn = nn.Sequential(*layers).to(self.dev)
        self.out = nn.Linear(prev, 1).to(self.dev)
        with torch.no_grad():
            self.out.weight.copy_((torch.randn(1, prev, generator=g) * math.sqrt(2.0 / (prev + 1))).to(self.dev))
            self.out.bias.zero_()
        self.wd = nn.Parameter((torch.randn(max(num_dense, 1), generator=g) * math.sqrt(2.0 / (num_dense + 1))).to(self.dev))
        self.bias = nn.Parameter(torch.zeros(1, device=self.dev))
        vc = sum(self.vocab[f] for f in self.cached)
        co, o = [], 0
        for f in self.cached:
            co.append(o)
            o += self.vocab[f]
        self.cache_off = torch.tensor(co or [0], dtype=torch.int64, device=self.dev)
        self.cache_cols = torch.tensor(self.cached or [0], dtype=torch.int64, device=self.dev)
        self.cache_tab = nn.Parameter(torch.zeros((max(vc, 1), embedding_dim + 1), device=self.dev))
        params = list(self.dnn.parameters()) + list(self.out.parameters()) + [self.wd, self.bias, self.cache_tab]
        # ---- flat fp32 master / grad / accumulator (one all-reduce, one optimizer pass)
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, device=self.dev)
        self.grad = torch.zeros(n, device=self.dev)
        self.acc = torch.full((n,), init_acc, device=self.dev)
        o = 0
        for p in params:
            k = p.numel()
            self.flat[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + k].view_as(p)
            p.grad = self.grad[o:o + k].view_as(p)
            o += k
        self.params = params
        self.nccl_calls = 0

    # ------------------------------------------------------------------ sparse exchange
    def _route(self, ids):
        """ids [B, F] -> (unique routing keys sorted by owner, inverse [B*ns], per-owner counts)"""
        gid = ids[:, self.srv_cols]                                   # [B, ns]
        owner = gid % self.W
        local = gid // self.W + self.row_off                          # row in the owner's concatenated slab
        key = owner * self.R + local
        uk, inv = torch.unique(key.reshape(-1), return_inverse=True)  # sorted => grouped by owner
        return uk, inv

    def _pull(self, uk):
        if self.W == 1:
            return self.table[uk], None
        owner = uk // self.R
        send_counts = torch.bincount(owner, minlength=self.W)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts)
        sc, rc = send_counts.tolist(), recv_counts.tolist()            # host sync: split sizes (as in every a2a embedding)
        req = torch.empty(sum(rc), dtype=torch.int64, device=self.dev)
        dist.all_to_all_single(req, (uk % self.R).contiguous(), output_split_sizes=rc, input_split_sizes=sc)
        rows = self.table[req]
        back = torch.empty((uk.numel(), self.D + 1), dtype=torch.float32, device=self.dev)
        dist.all_to_all_single(back, rows, output_split_sizes=sc, input_split_sizes=rc)
        self.nccl_calls += 3
        return back, (req, sc, rc)

    def _push_update(self, uk, g_u, route):
        if self.W == 1:
            rows, g = uk, g_u                                          # already unique and pre-reduced
        else:
            req, sc, rc = route
            recv = torch.empty((req.numel(), self.D + 1), dtype=torch.float32, device=self.dev)
            dist.all_to_all_single(recv, g_u.contiguous(), output_split_sizes=rc, input_split_sizes=sc)
            self.nccl_calls += 1
            rows, inv = torch.unique(req, return_inverse=True)         # combine the W sources
            g = torch.zeros((rows.numel(), self.D + 1), dtype=torch.float32, device=self.dev).index_add_(0, inv, recv)
        a = self.table_acc[rows] + g * g
        self.table_acc[rows] = a
        self.table[rows] -= self.slr * g / (a.sqrt() + self.eps)

    # ------------------------------------------------------------------ one step
    def step(self, ids, dense, labels):
        B, D = ids.shape[0], self.D
        self.grad.zero_()
        uk, inv = self._route(ids)
        rows, route = self._pull(uk)
        rows_u = rows.detach().requires_grad_(True)
        x = rows_u[inv].view(B, self.ns, D + 1)
        emb, lin = x[:, :, :D], x[:, :, D].sum(1)
        if self.nc:
            cx = self.cache_tab[ids[:, self.cache_cols] + self.cache_off]          # [B, nc, D+1]
            emb = torch.cat([emb, cx[:, :, :D]], dim=1)
            lin = lin + cx[:, :, D].sum(1)
        z = lin + self.bias
        if self.nd:
            z = z + dense @ self.wd[: self.nd]
        if self.model == "deepfm":
            s = emb.sum(1)
            z = z + 0.5 * (s * s - (emb * emb).sum(1)).sum(1)
        with torch.autocast("cuda", dtype=self.compute_dtype, enabled=self.compute_dtype != torch.float32):
            h = self.dnn(torch.cat([emb.reshape(B, -1), dense], dim=1))
            z = z + self.out(h).squeeze(-1).float()
        loss = F.binary_cross_entropy_with_logits(z, labels)
        loss.backward()
        self._push_update(uk, rows_u.grad, route)
        if self.W > 1:
            dist.all_reduce(self.grad)
            self.nccl_calls += 1
        self.acc.addcmul_(self.grad, self.grad)
        self.flat.addcdiv_(self.grad, self.acc.sqrt().add_(self.eps), value=-self.lr)
        return loss.detach()

    # ------------------------------------------------------------------ parity with the fused engine
    @torch.no_grad()
    def load_from_fused(self, m):
        """copy the dense weights of a ``FusedCTR`` (fp32 master copy) -- feature order there is
        [server features..., cached features..., dense]; padded columns are dropped"""
        assert m.D == self.D and m.nf == len(self.vocab) and m.server == self.server
        nf, Dp, D = m.nf, m.Dp, m.D
        dims = [m.K0p] + m.Hp
        lins = [l for l in self.dnn if isinstance(l, nn.Linear)]
        for l, lin in enumerate(lins):
            Wl = m.view("W%d" % l).view(m.Hp[l], dims[l])
            h = self.hidden[l]
            if l == 0:
                cols = torch.cat([torch.arange(f * Dp, f * Dp + D) for f in range(nf)] +
                                 [torch.arange(nf * Dp, nf * Dp + self.nd)]).to(self.dev)
                lin.weight.copy_(Wl[:h][:, cols])
            else:
                lin.weight.copy_(Wl[:h, : self.hidden[l - 1]])
            lin.bias.copy_(Wl[:h, dims[l] - 1])
        wo = m.view("wout")
        self.out.weight.copy_(wo[: self.hidden[-1]].view(1, -1))
        self.out.bias.copy_(wo[m.Hp[-1] - 1].view(1))
        if self.nd:
            self.wd[: self.nd].copy_(m.view("wd")[: self.nd])
        self.bias.copy_(m.view("bias"))


class HostPipeline:
    """End-to-end driver of the baseline: pinned host batch -> device (copy stream, double buffered), one
    ``step``, asynchronous D2H of the loss. Plain torch streams/events; mirrors what ``models.trainer._Pipeline``
    does for the engine so both arms pay the same host<->device traffic."""

    def __init__(self, model, batch, num_sparse, num_dense):
        self.m, dev = model, model.dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.dev = [dict(ids=torch.zeros((batch, num_sparse), dtype=torch.int64, device=dev),
                         dense=torch.zeros((batch, num_dense), dtype=torch.float32, device=dev),
                         labels=torch.zeros((batch,), dtype=torch.float32, device=dev)) for _ in range(2)]
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        self.h2d_bytes = batch * num_sparse * 8 + batch * num_dense * 4 + batch * 4
        self.d2h_bytes = 4

    def submit(self, ids_h, dense_h, labels_h):
        i = self.k & 1
        cur = torch.cuda.current_stream(self.m.dev)
        if self.k >= 2:
            self.copy_stream.wait_event(self.consumed[i])
        with torch.cuda.stream(self.copy_stream):
            d = self.dev[i]
            d["ids"].copy_(ids_h, non_blocking=True)
            d["dense"].copy_(dense_h, non_blocking=True)
            d["labels"].copy_(labels_h, non_blocking=True)
            self.copied[i].record(self.copy_stream)
        cur.wait_event(self.copied[i])
        if self.k >= 2:
            self.done[i].synchronize()
        loss = self.m.step(self.dev[i]["ids"], self.dev[i]["dense"], self.dev[i]["labels"])
        self.consumed[i].record(cur)
        self.loss_host[i].copy_(loss, non_blocking=True)
        self.done[i].record(cur)
        self.k += 1

    def last_loss(self):
        i = (self.k - 1) & 1
        self.done[i].synchronize()
        return float(self.loss_host[i])
