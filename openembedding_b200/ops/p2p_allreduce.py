"""Two-shot all-reduce (sum) over peer-mapped gradient buffers -- NVLink P2P, no NCCL.

Reference: the dense gradients go through Horovod / NCCL (``hvd.DistributedOptimizer(op=hvd.Sum)``,
test/benchmark/criteo_deepctr.py:262-263; examples/criteo_deepctr_network.py:54).

Kernels: ``exb_ar_*`` in ``csrc/cuda/dense_kernels.cu``. Every rank owns a cudaMalloc'd
gradient buffer and a flag block, both exported once with CUDA IPC; a call is ONE persistent
kernel (flag exchange, reduce-scatter by peer loads fused with the all-gather by peer stores,
flag exchange, optional fused Adagrad), graph-capturable.
"""
import ctypes

import torch

from .. import _native


def tensor_from_ptr(ptr, numel, device, dtype=torch.float32):
    """torch view of a raw device pointer (lifetime managed by the caller)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": {torch.float32: "<f4", torch.int32: "<i4"}[dtype],
                                  "data": (int(ptr), False), "version": 3, "strides": None}
    t = torch.as_tensor(h, device=device)
    t._exb_holder = h
    return t


class P2PAllReduce:
    """``P2PAllReduce(ctx, n)`` owns the peer-mapped gradient buffer (``.grad``: write the
    gradients there, no copies); ``P2PAllReduce(ctx, tensor)`` mirrors an existing flat
    tensor through two local copies. ``__call__(theta, accum, lr, eps)`` fuses the dense
    Adagrad step behind the reduction (one launch)."""

    def __init__(self, ctx, flat, ctas=0):
        import torch.distributed as dist
        self.ctx, self.W, self.rank = ctx, ctx.world, ctx.rank
        self.lib = _native.cuda()
        from ..models import fused_dense
        fused_dense._lib()          # prototypes of the dense kernels
        eng = ctx.backend.engine
        dev = ctx.device
        if isinstance(flat, int):
            n, self.flat = flat, None
        else:
            assert flat.is_cuda and flat.dtype == torch.float32
            n, self.flat = flat.numel(), flat
        assert n % 4 == 0
        self.n, self.dev = n, dev
        self.ctas = ctas
        self.buf_ptr = self.lib.exb_raw_alloc(eng.device_index, n * 4)
        self.flag_ptr = self.lib.exb_raw_alloc(eng.device_index, 2 << 20)
        if not self.buf_ptr or not self.flag_ptr:
            raise RuntimeError("exb_raw_alloc: " + self.lib.exb_cuda_last_error().decode())
        gathered = [None] * self.W
        dist.all_gather_object(gathered, (eng._export(self.buf_ptr), eng._export(self.flag_ptr)), group=ctx.group)
        self.bufs = (ctypes.c_uint64 * 8)()
        self.flags = (ctypes.c_uint64 * 8)()
        for r, (bh, fh) in enumerate(gathered):
            if r == self.rank:
                self.bufs[r], self.flags[r] = self.buf_ptr, self.flag_ptr
            else:
                self.bufs[r], self.flags[r] = eng._open(bh), eng._open(fh)
        self.local = tensor_from_ptr(self.buf_ptr, n, dev)
        self.grad = self.local
        self.local.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group=ctx.group)

    def __call__(self, opt_args=None):
        """opt_args: a ``models.fused_dense._DenseOptArgs`` whose ``grad`` is ``self.grad`` -- the dense
        optimizer step then runs inside the same kernel, behind the reduction."""
        st = torch.cuda.current_stream(self.dev).cuda_stream
        if self.flat is not None:
            self.local.copy_(self.flat, non_blocking=True)
        # flag block: [0, 32) flags, epoch word at +1024, status word at +2048, CTA counter at +3072
        rc = self.lib.exb_allreduce_adagrad(self.bufs, self.flags, self.flag_ptr + 1024, self.flag_ptr + 3072,
                                            self.flag_ptr + 2048, self.n, self.W, self.rank, self.ctas,
                                            ctypes.byref(opt_args) if opt_args is not None else None, st)
        if rc != 0:
            raise RuntimeError("exb_allreduce_adagrad: " + self.lib.exb_dense_last_error().decode())
        if self.flat is not None:
            self.flat.copy_(self.local, non_blocking=True)

    def phases_us(self):
        """in-kernel phase clock of the last call (CTA 0): wait for peers' gradients, reduce+broadcast,
        wait for peers' stores, optimizer"""
        t = tensor_from_ptr(self.flag_ptr + 4096, 14, self.dev, dtype=torch.int32).cpu().view(torch.int64).tolist()
        names = ["wait_ready", "reduce_bcast", "wait_landed", "optimizer"]
        out = {n: (t[i + 1] - t[i]) / 1e3 for i, n in enumerate(names) if t[i + 1] and t[i]}
        c = tensor_from_ptr(self.flag_ptr + 4096 + 128, 2 * 4 * 1200, self.dev, dtype=torch.int32).cpu().view(torch.int64).view(-1, 4)
        c = c[c[:, 0] > 0]
        if c.numel():
            t0 = int(c[:, 0].min())
            for k, nm in enumerate(["cta_start", "cta_ready", "cta_reduced", "cta_arrived"]):
                col = (c[:, k] - t0).double() / 1e3
                out[nm] = "min %.1f mean %.1f max %.1f (argmax cta %d of %d)" % (col.min(), col.mean(), col.max(), int(col.argmax()), c.shape[0])
        if t[5] and t[2]:
            out["last_cta_arrived_after_reduce"] = (t[5] - t[2]) / 1e3
            out["signal_issue"] = (t[6] - t[5]) / 1e3
        return out

    def status(self):
        t = tensor_from_ptr(self.flag_ptr + 2048, 1, self.dev, dtype=torch.int32)
        return int(t.item())
