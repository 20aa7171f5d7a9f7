"""Two-shot all-reduce (sum) over peer-mapped gradient buffers -- NVLink P2P, no NCCL.

Kernels: ``exb_ar_*`` in ``csrc/cuda/dense_kernels.cu``. Every rank owns a cudaMalloc'd
gradient mirror and a flag block, both exported once with CUDA IPC; a call is five stream
ordered launches (barrier, reduce-scatter by peer loads, barrier, all-gather by peer
stores, barrier) plus two local copies, all graph-capturable.
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
    def __init__(self, ctx, flat, ctas=64):
        import torch.distributed as dist
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.numel() % 4 == 0
        self.ctx, self.flat, self.W, self.rank = ctx, flat, ctx.world, ctx.rank
        self.lib = _native.cuda()
        from ..models import fused_dense
        fused_dense._lib()          # prototypes of the dense kernels
        eng = ctx.backend.engine
        dev = flat.device
        n = flat.numel()
        per = ((n + self.W - 1) // self.W + 3) // 4 * 4
        self.scratch = torch.empty(per, dtype=torch.float32, device=dev)
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
        torch.cuda.synchronize(dev)
        dist.barrier(group=ctx.group)

    def __call__(self):
        st = torch.cuda.current_stream(self.flat.device).cuda_stream
        self.local.copy_(self.flat, non_blocking=True)
        # flag block: [0, 32) flags, epoch word at +1024, status word at +2048
        rc = self.lib.exb_allreduce_sum(self.bufs, self.flags, self.flag_ptr + 1024, self.flag_ptr + 2048,
                                        self.scratch.data_ptr(), self.flat.numel(), self.W, self.rank, self.ctas, st)
        if rc != 0:
            raise RuntimeError("exb_allreduce_sum: " + self.lib.exb_dense_last_error().decode())
        self.flat.copy_(self.local, non_blocking=True)

    def status(self):
        t = tensor_from_ptr(self.flag_ptr + 2048, 1, self.flat.device, dtype=torch.int32)
        return int(t.item())
