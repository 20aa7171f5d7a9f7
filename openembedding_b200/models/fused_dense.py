"""Fused DeepFM / Wide&Deep training step on hand-written sm_100a kernels only.

Per step and per GPU (10 launches at world 1, captured in one CUDA graph by ``FusedTrainer``):

    pull + plan (peer loads, cp.async)   sparse_v2.cuh / sparse_kernels.cuh   (prefetched in the previous step's tail)
    prep                                 dense_kernels.cu   X32 -> A0 (bf16), FM sums, base logit
    3x GEMM fwd  (tcgen05, relu, ones)   gemm_tcgen05.cu
    head         (loss, dlogit, dZ_L)    dense_kernels.cu
    backward chain: 3x dX (relu mask / FM-fused fp32 embedding gradient) + 3x dW (MN-major operands, split-K,
                 TMA reduce-add) as ONE persistent kernel (exb_gemm_chain_kernel)
    cachegrad, push_update (P2P dispatch + combine + sparse optimizer; world > 1: the dense-gradient all-reduce
                 rides on its cross-GPU barriers),
    Adagrad / Adam / FTRL (flat) + bf16 weight refresh  ||  pull + plan of the NEXT batch on a side stream

No cuBLAS, no NCCL, no torch op on the step. Biases are folded into the GEMMs through a
constant "ones" column, so a layer is exactly one GEMM in each direction.

Model definition = DeepCTR's DeepFM / WDL as used by the reference benchmark
(test/benchmark/criteo_deepctr.py:243-282); see ``models/ctr.py`` for the eager version
of the same architecture (used as the numerical reference in tests).
"""
import ctypes
import os
import math
from ctypes import c_float, c_int, c_longlong, c_void_p

import torch

from .. import _native
from ..context import get_context
from ..ops import gemm as G
from ..utils import timers as _timers


class _PrepArgs(ctypes.Structure):
    _fields_ = [("X32", c_void_p), ("xs", c_longlong), ("A0", c_void_p), ("A0T", c_void_p), ("ids", c_void_p),
                ("ncols", c_int), ("dense", c_void_p), ("nd", c_int), ("cache_emb", c_void_p),
                ("cache_lin", c_void_p), ("cache_col", c_void_p), ("cache_off", c_void_p), ("nc", c_int),
                ("wd", c_void_p), ("bias", c_void_p), ("S", c_void_p), ("base", c_void_p), ("B", c_int),
                ("K0p", c_int), ("Dp", c_int), ("nf", c_int), ("ns", c_int), ("lin0", c_int), ("use_fm", c_int),
                ("loss", c_void_p), ("opt_step", c_void_p)]


class _HeadArgs(ctypes.Structure):
    _fields_ = [("H", c_void_p), ("Hp", c_int), ("ones_col", c_int), ("wout", c_void_p), ("base", c_void_p),
                ("labels", c_void_p), ("dlogit", c_void_p), ("loss", c_void_p), ("dZ", c_void_p), ("dZT", c_void_p),
                ("g_wout", c_void_p), ("g_wd", c_void_p), ("g_bias", c_void_p), ("dense", c_void_p), ("nd", c_int),
                ("G32", c_void_p), ("xs", c_longlong), ("lin0", c_int), ("ns", c_int), ("ids", c_void_p),
                ("ncols", c_int), ("cache_col", c_void_p), ("cache_off", c_void_p), ("nc", c_int),
                ("g_cache_lin", c_void_p), ("B", c_int), ("grad_scale", c_float)]


class _OptMat(ctypes.Structure):
    _fields_ = [("off", c_longlong), ("R", c_int), ("C", c_int), ("Wb", c_void_p), ("WTb", c_void_p)]


class _DenseOptArgs(ctypes.Structure):
    _fields_ = [("theta", c_void_p), ("accum", c_void_p), ("grad", c_void_p), ("n", c_longlong), ("flat_lo", c_longlong),
                ("lr", c_float), ("eps", c_float), ("nmat", c_int), ("zero_grad", c_int), ("mat", _OptMat * 4),
                ("kind", c_int), ("_pad", c_int), ("accum2", c_void_p), ("step", c_void_p), ("b1", c_float), ("b2", c_float),
                ("l1", c_float), ("l2", c_float), ("l2s", c_float), ("lrp", c_float), ("beta", c_float),
                ("c1", c_float), ("c2", c_float)]


def _r(x, m):
    return (x + m - 1) // m * m


_proto = False


def _lib():
    global _proto
    lib = _native.cuda()
    if not _proto:
        u64 = ctypes.c_uint64
        lib.exb_prep.restype = c_int
        lib.exb_prep.argtypes = [c_void_p, c_int, c_int, u64]
        lib.exb_head.restype = c_int
        lib.exb_head.argtypes = [c_void_p, c_int, u64]
        lib.exb_prep_args_size.restype = c_int
        lib.exb_head_args_size.restype = c_int
        lib.exb_cachegrad.restype = c_int
        lib.exb_cachegrad.argtypes = [u64, c_longlong, c_int, c_int, u64, c_int, u64, u64, c_int, u64, c_int, u64, u64, u64,
                                      u64]
        lib.exb_adagrad_flat.restype = c_int
        lib.exb_adagrad_flat.argtypes = [u64, u64, u64, c_longlong, c_float, c_float, u64]
        lib.exb_refresh_bf16.restype = c_int
        lib.exb_refresh_bf16.argtypes = [u64, u64, u64, c_int, c_int, u64]
        lib.exb_allreduce_adagrad.restype = c_int
        lib.exb_allreduce_adagrad.argtypes = [ctypes.POINTER(u64), ctypes.POINTER(u64), u64, u64, u64, c_longlong, c_int,
                                              c_int, c_int, c_void_p, u64]
        lib.exb_dense_opt.restype = c_int
        lib.exb_dense_opt.argtypes = [c_void_p, u64]
        assert lib.exb_dense_opt_args_size() == ctypes.sizeof(_DenseOptArgs), "DenseOptArgs ABI mismatch"
        lib.exb_dense_last_error.restype = ctypes.c_char_p
        assert lib.exb_prep_args_size() == ctypes.sizeof(_PrepArgs), "PrepArgs ABI mismatch"
        assert lib.exb_head_args_size() == ctypes.sizeof(_HeadArgs), "HeadArgs ABI mismatch"
        _proto = True
    return lib


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError("%s: %s" % (what, _lib().exb_dense_last_error().decode()))


class FusedCTR:
    """DeepFM (use_fm=True) or Wide&Deep (use_fm=False) with the whole step on own kernels."""

    def __init__(self, vocab_sizes, num_dense=13, embedding_dim=64, model="deepfm", batch=4096, hidden=None,
                 sparse_optimizer=None, cache_threshold=0, lr=0.001, initial_accumulator_value=0.1, eps=1e-7,
                 num_shards=None, dw_splits=8, seed=0, pack_linear=None, dense_optimizer=None):
        from .ctr import FusedEmbeddings
        ctx = get_context()
        if ctx.device.type != "cuda":
            raise RuntimeError("FusedCTR runs on the CUDA engine only (use models.ctr.CTRModel on CPU)")
        assert batch % 128 == 0, "the fused dense path needs batch % 128 == 0"
        self.ctx, self.dev, self.lib = ctx, ctx.device, _lib()
        self.model = model.lower()
        assert self.model in ("deepfm", "wdl")
        self.use_fm = self.model == "deepfm"
        self.B, self.nd, self.D = batch, num_dense, embedding_dim
        self.Dp = _r(embedding_dim, 4)
        self.vocab = list(vocab_sizes)
        self.nf = len(self.vocab)
        if hidden is None:
            hidden = (400, 400, 400) if self.use_fm else (512, 256, 128, 32)
        self.hidden = list(hidden)
        self.Hp = [_r(h + 1, 64) for h in self.hidden]
        self.lr, self.eps, self.dw_splits = lr, eps, int(os.environ.get("EXB_DW_SPLITS", dw_splits))
        self.cached = [f for f, v in enumerate(self.vocab) if 0 < v < cache_threshold]
        self.server = [f for f in range(self.nf) if f not in self.cached]
        self.ns, self.nc = len(self.server), len(self.cached)
        nf, Dp = self.nf, self.Dp
        self.K0p = _r(nf * Dp + num_dense + 1, 64)
        self.lin0 = self.K0p
        self.XS = _r(self.K0p + self.ns, 4)
        sparse_optimizer = sparse_optimizer or {"category": "adagrad"}
        zero = {"category": "constant", "value": 0.0}
        # pack_linear: the dim-D embedding and the dim-1 linear ("wide") weight of a sparse feature share ONE
        # server row of dim D+1 (split-row feature): half the lookups, unique ids, hash inserts, NVLink rows and
        # optimizer rows of the two-variables-per-feature layout of the reference benchmark
        # (criteo_deepctr.py:60-110). Same math per element; the checkpoint then holds one variable per feature.
        if pack_linear is None:
            pack_linear = os.environ.get("EXB_PACK_LINEAR", "1") != "0"
        self.pack_linear = bool(pack_linear)
        if self.pack_linear:
            specs = [{"vocab": self.vocab[f], "dim": embedding_dim + 1, "col": f, "initializer": zero} for f in self.server]
            self.sparse = FusedEmbeddings(specs, batch, sparse_optimizer, num_shards=num_shards, ncols=nf,
                                          feat_offsets=[j * Dp for j in range(self.ns)], io_stride=self.XS,
                                          feat_offsets2=[self.lin0 + j for j in range(self.ns)],
                                          feat_split=[embedding_dim] * self.ns)
        else:
            specs = [{"vocab": self.vocab[f], "dim": embedding_dim, "col": f, "initializer": zero} for f in self.server]
            specs += [{"vocab": self.vocab[f], "dim": 1, "col": f, "initializer": zero} for f in self.server]
            offs = [j * Dp for j in range(self.ns)] + [self.lin0 + j for j in range(self.ns)]
            self.sparse = FusedEmbeddings(specs, batch, sparse_optimizer, num_shards=num_shards, ncols=nf,
                                          feat_offsets=offs, io_stride=self.XS)
        self.group = self.sparse.group
        dev = self.dev
        f32, bf16 = torch.float32, torch.bfloat16
        # ---- flat parameter buffer
        segs, off = {}, 0
        dims = [self.K0p] + self.Hp
        for l in range(len(self.hidden)):
            n = self.Hp[l] * dims[l]
            segs["W%d" % l] = (off, n)
            off = _r(off + n, 4)
        L = len(self.hidden)
        for name, n in (("wout", self.Hp[-1]), ("wd", max(num_dense, 1)), ("bias", 1)):
            segs[name] = (off, n)
            off = _r(off + n, 4)
        vc = sum(self.vocab[f] for f in self.cached)
        self.cache_rows = vc
        segs["cache_emb"] = (off, vc * Dp)
        off = _r(off + vc * Dp, 4)
        segs["cache_lin"] = (off, vc)
        off = _r(off + max(vc, 1), 4)
        self.segs, self.n_theta = segs, off
        self.theta = torch.zeros(off, dtype=f32, device=dev)
        # dense optimizer (tf.keras semantics): {"category": "adagrad" | "adam" | "ftrl", ...}; default Adagrad(lr)
        from ..config import normalize_optimizer
        dopt = normalize_optimizer(dense_optimizer or {"category": "adagrad", "learning_rate": lr,
                                                        "initial_accumulator_value": initial_accumulator_value, "epsilon": eps})
        if dopt["category"] not in ("adagrad", "adam", "ftrl"):
            raise ValueError("fused dense optimizer: adagrad, adam or ftrl")
        self.dense_opt = dopt
        self.lr = lr = float(dopt["learning_rate"])
        acc0 = float(dopt.get("initial_accumulator_value", 0.0)) if dopt["category"] in ("adagrad", "ftrl") else 0.0
        self.accum = torch.full((off,), acc0, dtype=f32, device=dev)
        self.accum2 = torch.zeros(off if dopt["category"] != "adagrad" else 4, dtype=f32, device=dev)
        self.opt_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ar = None
        self._rider = False
        self.overlap = os.environ.get("EXB_OVERLAP", "0") == "1"
        self.late_cachegrad = os.environ.get("EXB_LATE_CACHEGRAD", "0") == "1"     # measured: no gain (0.2260 vs 0.2240), the tail pull owns the SMs
        if ctx.world > 1:     # gradients are produced straight into the peer-mapped all-reduce buffer
            from ..ops.p2p_allreduce import P2PAllReduce
            self._ar = P2PAllReduce(ctx, off)
            self.gtheta = self._ar.grad
            # the reduction rides on the sparse push kernel's cross-GPU barriers instead of being a kernel with two
            # barriers of its own (EXB_AR_RIDER=0: separate exb_ar_fused_kernel launch)
            self._rider = (os.environ.get("EXB_AR_RIDER", "1") != "0" and not self.overlap
                           and hasattr(self.group, "set_dense_reduce"))
            if self._rider:
                self.group.set_dense_reduce(self._ar.bufs, off)
        else:
            self.gtheta = torch.zeros(off, dtype=f32, device=dev)
        gen = torch.Generator(device="cpu").manual_seed(seed)
        fan_in = [nf * embedding_dim + num_dense] + self.hidden
        for l in range(L):
            W = self.view("W%d" % l).view(self.Hp[l], dims[l])
            h_out = self.hidden[l]
            std = math.sqrt(2.0 / (fan_in[l] + h_out))            # glorot normal (DeepCTR DNN default)
            real_in = nf * Dp + num_dense if l == 0 else self.hidden[l - 1]
            blk = torch.randn(h_out, real_in, generator=gen) * std
            if l == 0 and Dp != embedding_dim:                     # zero the weights that face pad columns
                m = torch.ones(nf, Dp)
                m[:, embedding_dim:] = 0
                blk[:, :nf * Dp] *= m.reshape(-1)
            W[:h_out, :real_in] = blk.to(dev)
        self.view("wout")[: self.hidden[-1]] = (torch.randn(self.hidden[-1], generator=gen)
                                                 * math.sqrt(2.0 / (self.hidden[-1] + 1))).to(dev)
        if num_dense:
            self.view("wd")[:num_dense] = (torch.randn(num_dense, generator=gen) * math.sqrt(2.0 / (num_dense + 1))).to(dev)
        # ---- bf16 K-major weight copies
        self.Wb = [torch.zeros(self.Hp[l], dims[l], dtype=bf16, device=dev) for l in range(L)]
        self.WTb = [torch.zeros(dims[l], self.Hp[l], dtype=bf16, device=dev) for l in range(L)]
        # ---- activations
        B = batch
        self.X32 = torch.zeros(B, self.XS, dtype=f32, device=dev)
        self.G32 = torch.zeros(B, self.XS, dtype=f32, device=dev)
        self.A0 = torch.zeros(B, self.K0p, dtype=bf16, device=dev)
        self.A0T = torch.zeros(self.K0p, B, dtype=bf16, device=dev)
        self.H = [torch.zeros(B, hp, dtype=bf16, device=dev) for hp in self.Hp]
        self.HT = [torch.zeros(hp, B, dtype=bf16, device=dev) for hp in self.Hp]
        self.dZ = [torch.zeros(B, hp, dtype=bf16, device=dev) for hp in self.Hp]
        self.dZT = [torch.zeros(hp, B, dtype=bf16, device=dev) for hp in self.Hp]
        self.S = torch.zeros(B, Dp, dtype=f32, device=dev)
        self.base = torch.zeros(B, dtype=f32, device=dev)
        self.dlogit = torch.zeros(B, dtype=f32, device=dev)
        self.loss = torch.zeros(1, dtype=f32, device=dev)
        offs_c, o = [], 0
        for f in self.cached:
            offs_c.append(o)
            o += self.vocab[f]
        self.cache_col = torch.tensor(self.cached or [0], dtype=torch.int32, device=dev)
        self.cache_off = torch.tensor(offs_c or [0], dtype=torch.int64, device=dev)
        self.cache_vocab = torch.tensor([self.vocab[f] for f in self.cached] or [0], dtype=torch.int32, device=dev)
        # push+update runs on a second stream next to the dW GEMMs / dense optimizer (fork after dX1, join at step end)
        self.overlap = os.environ.get("EXB_OVERLAP", "0") == "1"
        # weight-gradient GEMMs read the batch-major activations as MN-major operands: no A0^T / H^T / dZ^T copies
        self.mn_major = os.environ.get("EXB_MN_MAJOR", "1") != "0"
        self._s2 = torch.cuda.Stream(device=dev)
        self._ev_fork, self._ev_join, self._ev_plan = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        assert L <= 4, "the fused optimizer kernel takes at most 4 weight matrices"
        oa = _DenseOptArgs()
        oa.theta, oa.accum, oa.grad = self.theta.data_ptr(), self.accum.data_ptr(), self.gtheta.data_ptr()
        oa.n, oa.flat_lo, oa.lr, oa.eps = self.n_theta, segs["wout"][0], self.lr, float(self.dense_opt.get("epsilon", self.eps))
        oa.nmat, oa.zero_grad = L, 1
        d = self.dense_opt
        oa.kind = {"adagrad": 0, "adam": 1, "ftrl": 2}[d["category"]]
        oa.accum2, oa.step = self.accum2.data_ptr(), self.opt_step.data_ptr()
        oa.b1, oa.b2 = float(d.get("beta_1", 0.9)), float(d.get("beta_2", 0.999))
        oa.l1, oa.l2 = float(d.get("l1_regularization_strength", 0.0)), float(d.get("l2_regularization_strength", 0.0))
        oa.l2s, oa.lrp = float(d.get("l2_shrinkage_regularization_strength", 0.0)), float(d.get("learning_rate_power", -0.5))
        oa.beta = float(d.get("beta", 0.0))
        for l in range(L):
            oa.mat[l].off, oa.mat[l].R, oa.mat[l].C = segs["W%d" % l][0], self.Hp[l], dims[l]
            oa.mat[l].Wb, oa.mat[l].WTb = self.Wb[l].data_ptr(), self.WTb[l].data_ptr()
        self._opt_args = oa
        self._grad_dirty = False
        # persistent GEMM chains: forward (fwd1 -> ... -> fwdL) and backward (dX / dW of every layer) in ONE launch
        # each (csrc/cuda/gemm_tcgen05.cu: exb_gemm_chain_kernel). EXB_GEMM_CHAIN=0: one launch per GEMM.
        mode = os.environ.get("EXB_GEMM_CHAIN", "bwd")          # "0" | "bwd" | "1" (forward and backward)
        self.use_chain = mode != "0" and self.mn_major and L <= 4
        self.chain_fwd = self.use_chain and mode == "1"
        self.fwd_chain = self.bwd_chain = None
        if self.use_chain:
            fd, src = [], self.A0
            for l in range(L):
                fd.append(G.chain_nt(src, self.Wb[l], B, self.Hp[l], dims[l], self.H[l], mode=G.EPI_FWD, relu=True,
                                     ones_col=self.Hp[l] - 1, dep=l - 1))
                src = self.H[l]
            self.fwd_chain = G.GemmChain(fd, dev)
            bd, prod = [], -1          # prod: index (in the chain) of the GEMM that produced dZ[l]
            for l in range(L - 1, -1, -1):
                gW = self.gview("W%d" % l).view(self.Hp[l], dims[l])
                if l > 0:
                    bd.append(G.chain_nt(self.dZ[l], self.WTb[l], B, self.Hp[l - 1], self.Hp[l], self.dZ[l - 1], mode=G.EPI_DX,
                                         ones_col=self.Hp[l - 1] - 1, mask=self.H[l - 1], dep=prod))
                else:
                    bd.append(G.chain_nt(self.dZ[0], self.WTb[0], B, self.K0p, self.Hp[0], self.G32, mode=G.EPI_DX_FM,
                                         dlogit=self.dlogit, S=self.S, emb=self.X32,
                                         fm_cols=self.nf * self.Dp if self.use_fm else 0, D=self.Dp, dep=prod))
                nxt = len(bd) - 1
                bd.append(G.chain_tn(self.dZ[l], self.A0 if l == 0 else self.H[l - 1], self.Hp[l], dims[l], B, gW,
                                     splits=self.dw_splits, dep=prod))
                prod = nxt
            self.bwd_chain = G.GemmChain(bd, dev)
        self.refresh_weights()
        torch.cuda.synchronize(dev)

    # ---- helpers
    def view(self, name):
        o, n = self.segs[name]
        return self.theta[o:o + n]

    def gview(self, name):
        o, n = self.segs[name]
        return self.gtheta[o:o + n]

    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    _trace = None          # list of (name, event) when stage tracing is on (tools/mp_timeline.py)

    def _mark(self, name):
        _timers.nvtx_mark(name)              # EXB_NVTX=1: stage boundaries on the nsys / ncu timeline
        if self._trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.dev))
            self._trace.append((name, ev))

    def refresh_weights(self):
        dims = [self.K0p] + self.Hp
        for l in range(len(self.hidden)):
            _ck(self.lib.exb_refresh_bf16(self.view("W%d" % l).data_ptr(), self.Wb[l].data_ptr(),
                                          self.WTb[l].data_ptr(), self.Hp[l], dims[l], self._st()), "refresh_bf16")

    # ---- one training step (all launches on the current stream)
    def forward_backward(self, ids, dense, labels, update=True, next_ids=None, pulled=False):
        """One training step. ``next_ids``: ids of the NEXT batch (a device tensor that stays unchanged until that
        batch has been trained): after this batch's push+update the rows AND the de-duplication plan of the next
        batch are pulled on a side stream, next to the dense all-reduce / optimizer of this step -- the prefetch of
        the reference's ``pulling`` (exb.py:645-691; parked pulls, EmbeddingPullOperator.cpp:117-145: a pull of batch
        k+1 may only see the tables after update k, which the stream order guarantees). ``pulled=True``: the caller
        passes the batch that the previous step prefetched this way: X32 and the plan are ready, no pull up front."""
        B, L, lib, st = self.B, len(self.hidden), self.lib, self._st()
        assert ids.shape == (B, self.nf) and ids.dtype == torch.int64 and ids.is_contiguous()
        if self._grad_dirty:          # the optimizer kernel clears the gradients it consumed; a call with
            self.gtheta.zero_()       # update=False leaves them behind
            self._grad_dirty = False
        self._mark("start")
        g = self.group
        v2 = getattr(g, "v2", False) and update
        if pulled:
            if v2:
                g._armed[0] = (g._key(ids), "pull")    # rows + plan of this batch came with the previous step's tail
        elif v2:
            g.pull(ids, out=self.X32, train=True)      # gather + plan of the batch in one launch
        else:
            g.pull(ids, out=self.X32)
        self._mark("pull")
        tn = self.mn_major
        pa = _PrepArgs(self.X32.data_ptr(), self.XS, self.A0.data_ptr(), 0 if tn else self.A0T.data_ptr(), ids.data_ptr(), self.nf,
                       dense.data_ptr(), self.nd, self.view("cache_emb").data_ptr(), self.view("cache_lin").data_ptr(),
                       self.cache_col.data_ptr(), self.cache_off.data_ptr(), self.nc, self.view("wd").data_ptr(),
                       self.view("bias").data_ptr(), self.S.data_ptr(), self.base.data_ptr(), B, self.K0p, self.Dp,
                       self.nf, self.ns, self.lin0, int(self.use_fm), self.loss.data_ptr(),
                       self.opt_step.data_ptr() if update else 0)
        _ck(lib.exb_prep(ctypes.byref(pa), B, self.Dp, st), "prep")
        self._mark("prep")
        dims = [self.K0p] + self.Hp
        src = self.A0
        if self.chain_fwd:
            self.fwd_chain.launch(st)
        else:
            for l in range(L):
                G.gemm_nt(src, self.Wb[l], B, self.Hp[l], dims[l], self.H[l], mode=G.EPI_FWD, relu=True,
                          ones_col=self.Hp[l] - 1, outT=self.HT[l] if (l < L - 1 and not tn) else None, stream=st)
                src = self.H[l]
        self._mark("fwd_gemm")
        row_head = tn and self.Hp[-1] <= 512       # merged row-wise head; cachegrad then owns the cached linear grads
        ha = _HeadArgs(self.H[-1].data_ptr(), self.Hp[-1], self.Hp[-1] - 1, self.view("wout").data_ptr(),
                       self.base.data_ptr(), labels.data_ptr(), self.dlogit.data_ptr(), self.loss.data_ptr(),
                       self.dZ[-1].data_ptr(), 0 if tn else self.dZT[-1].data_ptr(), self.gview("wout").data_ptr(),
                       self.gview("wd").data_ptr(), self.gview("bias").data_ptr(), dense.data_ptr(), self.nd,
                       self.G32.data_ptr(), self.XS, self.lin0, self.ns, ids.data_ptr(), self.nf,
                       self.cache_col.data_ptr(), self.cache_off.data_ptr(), self.nc,
                       0 if row_head else self.gview("cache_lin").data_ptr(), B, 1.0 / B)
        _ck(lib.exb_head(ctypes.byref(ha), B, st), "head")
        self._mark("head")
        if self.use_chain:
            self.bwd_chain.launch(st)          # dX and dW of every layer: one persistent launch
        else:
            for l in range(L - 1, 0, -1):      # dZ_{l-1} = (dZ_l @ W_l) * relu'(H_{l-1})
                G.gemm_nt(self.dZ[l], self.WTb[l], B, self.Hp[l - 1], self.Hp[l], self.dZ[l - 1], mode=G.EPI_DX,
                          ones_col=self.Hp[l - 1] - 1, outT=None if tn else self.dZT[l - 1], mask=self.H[l - 1], stream=st)
            G.gemm_nt(self.dZ[0], self.WTb[0], B, self.K0p, self.Hp[0], self.G32, mode=G.EPI_DX_FM, dlogit=self.dlogit,
                      S=self.S, emb=self.X32, fm_cols=self.nf * self.Dp if self.use_fm else 0, D=self.Dp, stream=st)
        self._mark("dx_gemm")
        forked = update and self.overlap
        if forked:
            cur = torch.cuda.current_stream(self.dev)
            self._ev_fork.record(cur)
            with torch.cuda.stream(self._s2):
                self._s2.wait_event(self._ev_fork)
                self.group.push_update(ids, self.G32)
                self._ev_join.record(self._s2)
        for l in range(L if not self.use_chain else 0):                 # dW_l = dZ_l^T @ H_{l-1}
            prevT = self.A0T if l == 0 else self.HT[l - 1]
            gW = self.gview("W%d" % l).view(self.Hp[l], dims[l])
            if tn:
                G.gemm_tn(self.dZ[l], self.A0 if l == 0 else self.H[l - 1], self.Hp[l], dims[l], B, gW,
                          splits=self.dw_splits, stream=st)
                continue
            G.gemm_nt(self.dZT[l], prevT, self.Hp[l], dims[l], B, gW, mode=G.EPI_DW, splits=self.dw_splits, stream=st)
        def cachegrad():
            _ck(lib.exb_cachegrad(self.G32.data_ptr(), self.XS, self.ns * self.Dp, self.Dp, ids.data_ptr(), self.nf,
                                  self.cache_col.data_ptr(), self.cache_off.data_ptr(), self.nc,
                                  self.gview("cache_emb").data_ptr(), B, self.dlogit.data_ptr(),
                                  self.gview("cache_lin").data_ptr() if row_head else 0,
                                  self.cache_vocab.data_ptr(), st), "cachegrad")
        tail = update and next_ids is not None and not forked
        # one GPU: the gradients of the replicated tables are only needed by the dense optimizer, so their kernel moves
        # behind push+update, next to the tail pull (several GPUs: the push kernel all-reduces them, they come first)
        late_cg = bool(self.nc) and tail and self._ar is None and self.late_cachegrad
        if self.nc and not late_cg:
            cachegrad()
        self._mark("dw_gemm+cachegrad")
        if update:
            if not forked:
                self.group.push_update(ids, self.G32)
                self._mark("push_update")
            if tail:      # next batch: rows into X32 (free since the dX1 GEMM) + plan, beside the dense optimizer
                cur = torch.cuda.current_stream(self.dev)
                self._ev_fork.record(cur)
                self._s2.wait_event(self._ev_fork)
                with torch.cuda.stream(self._s2):
                    g.pull(next_ids, out=self.X32, train=v2)
                    self._ev_plan.record(self._s2)
            # Adagrad + bf16 weight refresh + gradient clearing: one kernel (world > 1: behind the all-reduce,
            # in the same kernel)
            # world > 1: the all-reduce runs on one CTA per SM (every CTA polls peer flags); the optimizer kernel
            # is chained behind it with a programmatic dependent launch instead of sharing its grid
            if self._ar is not None and not self._rider:
                self._ar()
                self._mark("allreduce")
            if late_cg:
                cachegrad()
            _ck(lib.exb_dense_opt(ctypes.byref(self._opt_args), st), "dense_opt")
            self._mark("optimizer")
            if forked:
                torch.cuda.current_stream(self.dev).wait_event(self._ev_join)
                self._mark("join(push_update)")
            if tail:
                torch.cuda.current_stream(self.dev).wait_event(self._ev_plan)
                self._mark("join(prefetch pull)")
        else:
            self._grad_dirty = True
        return self.loss.view(())

    def warmup(self, ids, dense, labels):
        """Launch every kernel of a training step once WITHOUT changing a parameter: forward + backward of the batch
        with the gradients discarded, the planned pull and push+update of a zero-row batch (all phases and barriers,
        no row; world > 1: the ride-along all-reduce sums zero gradients), the dense optimizer on a snapshot that is
        restored afterwards. Used before a CUDA-graph capture -- lazy kernel loading and allocator growth may not
        happen inside one. Collective when world > 1."""
        lib, st = self.lib, self._st()
        keep = [t.clone() for t in (self.theta, self.accum, self.accum2, self.opt_step)]
        self.forward_backward(ids, dense, labels, update=False)
        g = self.group
        none_ids, none_g = ids[:0], self.G32[:0]
        if getattr(g, "v2", False):
            g.pull(none_ids, out=self.X32, train=True)
        self.gtheta.zero_()
        g.push_update(none_ids, none_g)
        if self._ar is not None and not self._rider:
            self._ar()
        _ck(lib.exb_dense_opt(ctypes.byref(self._opt_args), st), "dense_opt")
        for t, k in zip((self.theta, self.accum, self.accum2, self.opt_step), keep):
            t.copy_(k)
        self.gtheta.zero_()
        self._grad_dirty = False
        self.refresh_weights()

    def kernels_per_step(self):
        """launches of our own kernels in one training step"""
        L = len(self.hidden)
        prep = 1 if (self.mn_major and 128 % self.Dp == 0) else 2
        head = 1 if (self.mn_major and self.Hp[-1] <= 512) else 2
        gemms = (1 if self.chain_fwd else L) + (1 if self.use_chain else 2 * L)   # persistent chains: fwd, bwd
        n = 1 + prep + gemms + head + (1 if self.nc else 0) + 1 + 1      # pull prep GEMMs head cache push optimizer
        return n + (1 if self._ar is not None and not self._rider else 0)

    # ---- fp32 torch reference of the dense math on the current X32 (tests)
    def reference(self, ids, dense, labels):
        """returns (loss, grads dict) computed with torch autograd in fp32 from the same
        parameters and the same pulled embeddings (self.X32 after a forward)."""
        B, nf, Dp, L = self.B, self.nf, self.Dp, len(self.hidden)
        theta = self.theta.detach().clone().requires_grad_(True)
        X = self.X32.detach().clone()
        emb = X[:, :nf * Dp].clone()
        if self.nc:
            ce = theta[self.segs["cache_emb"][0]:self.segs["cache_emb"][0] + self.cache_rows * Dp].view(-1, Dp)
            cid = ids[:, self.cache_col.long()] + self.cache_off
            emb = torch.cat([emb[:, :self.ns * Dp], ce[cid].reshape(B, -1)], dim=1)
        emb = emb.detach().requires_grad_(True) if not self.nc else emb
        emb_leaf = X[:, :self.ns * Dp].clone().requires_grad_(True)
        if self.nc:
            emb = torch.cat([emb_leaf, ce[cid].reshape(B, -1)], dim=1)
        else:
            emb = emb_leaf
        lin_leaf = X[:, self.lin0:self.lin0 + self.ns].clone().requires_grad_(True)
        lin = lin_leaf.sum(1)
        if self.nc:
            cl = theta[self.segs["cache_lin"][0]:self.segs["cache_lin"][0] + self.cache_rows]
            lin = lin + cl[cid].sum(1)
        wd = theta[self.segs["wd"][0]:self.segs["wd"][0] + self.nd]
        bias = theta[self.segs["bias"][0]]
        z = lin + dense @ wd + bias
        if self.use_fm:
            e = emb.view(B, nf, Dp)
            s = e.sum(1)
            z = z + 0.5 * (s * s - (e * e).sum(1)).sum(1)
        ones = torch.ones(B, 1, device=self.dev)
        dims = [self.K0p] + self.Hp
        pad0 = self.K0p - 1 - nf * Dp - self.nd
        h = torch.cat([emb, dense, torch.zeros(B, pad0, device=self.dev), ones], dim=1)
        for l in range(L):
            o, n = self.segs["W%d" % l]
            W = theta[o:o + n].view(self.Hp[l], dims[l])
            h = torch.relu(h.to(torch.bfloat16).float() @ W.to(torch.bfloat16).float().t())
            h = torch.cat([h[:, :-1], ones], dim=1)
        o, n = self.segs["wout"]
        z = z + h.to(torch.bfloat16).float() @ theta[o:o + n]
        loss = torch.nn.functional.binary_cross_entropy_with_logits(z, labels)
        loss.backward()
        return loss.detach(), {"theta": theta.grad, "emb": emb_leaf.grad, "lin": lin_leaf.grad}


class FusedTrainer:
    """CUDA-graph driver for ``FusedCTR`` with the same interface as ``models.trainer.Trainer``.

    ``step(ids, dense, labels, next_ids=...)``: with ``next_ids`` (the device ids of the batch that will be passed
    as ``ids`` to the NEXT call) the de-duplication plan of the next batch is built inside this step on a side
    stream -- the prefetch of the reference's ``pulling`` (exb.py:645-691).

    Before the first capture every kernel of the step is launched eagerly twice through ``FusedCTR.warmup`` (lazy
    kernel loading and allocator growth may not happen inside a capture); the warm-up changes no parameter, so the
    graph-driven trajectory is the eager one from the first step on."""

    supports_prefetch = True
    want_prefetch = True         # ``make_pipeline`` runs one batch ahead: the next batch's pull overlaps this step's tail
    supports_stable_inputs = True
    MAX_STABLE_GRAPHS = 64

    def __init__(self, model, use_graph=True):
        self.m, self.ctx = model, model.ctx
        self.device, self.world = model.dev, model.ctx.world
        self.use_graph = use_graph
        self._graphs, self._static = {}, None      # (pull up front?, prefetch pull at the tail?[, input addresses]) -> CUDAGraph
        self._stable = {}                          # stable-input graph key -> the caller's tensors (kept alive)
        self.graph = None
        self._ar = model._ar
        self._x32_key = None         # key of the batch whose rows + plan the last step prefetched into X32
        self._warm = False

    def step(self, ids, dense, labels, next_ids=None, stable=False):
        """``stable=True``: the caller keeps the four input tensors alive at fixed addresses and refills them in place
        (an input pipeline's device buffers, a resident pool of batches). In the steady state (this batch was
        prefetched, the next one is announced) the graph is then captured directly on those tensors -- one graph per
        distinct (ids, dense, labels, next_ids) address tuple, at most MAX_STABLE_GRAPHS -- instead of copying the
        inputs into the trainer's own static buffers first (four device copies per step)."""
        g = self.m.group
        v2 = getattr(g, "v2", False)
        pulled = self._x32_key is not None and self._x32_key == g._key(ids)
        tail = next_ids is not None
        if not self.use_graph:
            if v2 and not pulled and self._x32_key is not None:
                g.reset_slot(0)                     # a prefetched batch that is not the one trained now
            loss = self.m.forward_backward(ids, dense, labels, next_ids=next_ids, pulled=pulled)
            self._x32_key = g._key(next_ids) if tail else None
            self.ctx.step_done()
            return loss
        if self._static is None:
            self._static = {"ids": ids.clone(), "dense": dense.clone(), "labels": labels.clone(), "next_ids": ids.clone()}
        s = self._static
        key = (not pulled, tail)
        if stable and pulled and tail and self._warm:
            skey = key + (ids.data_ptr(), dense.data_ptr(), labels.data_ptr(), next_ids.data_ptr())
            if skey in self._graphs or len(self._stable) < self.MAX_STABLE_GRAPHS:
                key = skey
                s = self._stable.setdefault(skey, {"ids": ids, "dense": dense, "labels": labels, "next_ids": next_ids})
        if s is self._static:
            if ids.data_ptr() != s["ids"].data_ptr():
                s["ids"].copy_(ids, non_blocking=True)
                s["dense"].copy_(dense, non_blocking=True)
                s["labels"].copy_(labels, non_blocking=True)
            if tail:
                s["next_ids"].copy_(next_ids, non_blocking=True)
        if v2 and not pulled and self._x32_key is not None:
            g.reset_slot(0)                         # drop the prefetched plan: this is a different batch
        gr = self._graphs.get(key)
        if gr is None:
            gr = self._capture(key, s)
        gr.replay()
        # replays bypass the plan's python bookkeeping: if a batch was prefetched the current slot is armed on the
        # device under a key no tensor can match -- any eager use of the plan re-plans
        if v2:
            g._armed = [(("graph", id(self)), "pull") if tail else None, None]
        self._x32_key = g._key(next_ids) if tail else None
        self.graph = gr
        self.ctx.step_done()
        return self._static["loss"]

    def _capture(self, key, s):
        head, tail = key[:2]
        g = self.m.group
        if not self._warm:          # allocator / lazy init warm-up, outside any capture
            assert head, "the first step of a trainer always pulls up front"
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):      # every kernel of the step runs, no parameter changes (FusedCTR.warmup)
                    self.m.warmup(s["ids"], s["dense"], s["labels"])
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                self.ctx.barrier()
            self._warm = True
        torch.cuda.synchronize(self.device)
        saved = list(getattr(g, "_armed", [None, None]))
        if getattr(g, "v2", False):
            g._armed = [None, None]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            loss = self.m.forward_backward(s["ids"], s["dense"], s["labels"], next_ids=s["next_ids"] if tail else None,
                                           pulled=not head)
        self._static["loss"] = loss          # the model's own loss buffer: the same tensor in every variant
        if getattr(g, "v2", False):
            g._armed = saved            # capture only recorded launches: the device-side slots are untouched
        self._graphs[key] = gr
        return gr

    def make_pipeline(self, batch, num_sparse, num_dense):
        from .trainer import _Pipeline
        return _Pipeline(self, batch, num_sparse, num_dense)
