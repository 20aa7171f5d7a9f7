"""Dense-gradient all-reduce (sum) for the data-parallel part of the model.

The reference delegates this to Horovod/NCCL (`hvd.DistributedOptimizer(op=hvd.Sum)`,
examples/criteo_deepctr_network.py:54). Here:

* ``mode="p2p"``  -- the product path: one-shot / two-shot reduction kernel over
  peer-mapped buffers (``exb_ar_fused_kernel`` in ``csrc/cuda/dense_kernels.cu``), no NCCL on the step; the fused
  trainer goes one step further and sums the gradients inside the sparse push kernel (``dense_reduce_gather`` /
  ``dense_reduce_scatter`` in ``csrc/cuda/sparse_kernels.cuh``);
* ``mode="nccl"`` -- the baseline / cross-check path (``torch.distributed.all_reduce``).
"""
import os

import torch


def make_allreduce(ctx, flat_grad, mode="auto"):
    """returns a zero-arg callable that sums `flat_grad` in place across ranks, or None to
    let the caller fall back to torch.distributed.all_reduce"""
    if mode == "auto":
        mode = os.environ.get("EXB_ALLREDUCE", "p2p")
    if mode == "nccl" or ctx.world == 1:
        return None
    try:
        from ..ops.p2p_allreduce import P2PAllReduce
    except ImportError:
        return None
    return P2PAllReduce(ctx, flat_grad)
