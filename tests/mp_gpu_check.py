"""Multi-process / multi-GPU correctness check (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mp_gpu_check.py

Checks, over real CUDA-IPC peer mappings: fused pull / push+update vs the CPU oracle,
P2P all-reduce vs a torch sum, collective checkpoint save + re-sharded load.
"""
import ctypes
import os
import shutil
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import openembedding_b200 as oe
    from openembedding_b200 import _native
    from openembedding_b200.config import initializer_params, mix_seed, optimizer_params
    from openembedding_b200.context import get_context
    oe.flags.device = "cuda"
    ctx = get_context()
    dev = ctx.device
    be = ctx.backend
    specs = [(64, 100000, False), (16, 2 ** 63, True), (1, 5000, False), (9, 77, False)]
    init = {"category": "uniform", "minval": -0.5, "maxval": 0.5, "seed": 0}
    opt = {"category": "adagrad", "learning_rate": 0.1}
    metas = []
    for dim, vocab, is_hash in specs:
        st = ctx.create_storage(None)
        m = ctx.create_variable(st, vocab, dim, "float32")
        ctx.set_initializer(m, init)
        ctx.set_optimizer(m, opt)
        metas.append(m)
    B = 512
    plan = be.make_group(metas, B)
    sl = plan.feature_slices()
    lib = _native.core()
    oracles = []
    for vid, (dim, vocab, is_hash) in enumerate(specs):
        h = lib.exb_var_create(0x104, dim, 0 if is_hash else vocab, 0, 1, 1 if is_hash else 0)
        k, p, seed = initializer_params(init)
        lib.exb_var_set_initializer(h, k, p[0], p[1], p[2], mix_seed(seed, metas[vid].variable_id))
        ok, op = optimizer_params(opt)
        lib.exb_var_set_optimizer(h, ok, (ctypes.c_double * 8)(*op), 8)
        oracles.append(h)
    worst = 0.0
    g = torch.Generator().manual_seed(100 + rank)
    for step in range(4):
        cols = []
        for dim, vocab, is_hash in specs:
            c = torch.randint(0, min(vocab, 400), (B,), generator=g)
            if is_hash:
                c = c * 1000003 + 7
            cols.append(c)
        ids = torch.stack(cols, 1).contiguous().to(dev)
        grads = torch.randn(B, plan.io_stride, generator=g).to(dev)
        out = plan.pull(ids)
        torch.cuda.synchronize()
        for f, (dim, vocab, is_hash) in enumerate(specs):
            k = np.ascontiguousarray(ids[:, f].cpu().numpy(), dtype=np.uint64)
            want = np.empty((B, dim), dtype=np.float32)
            lib.exb_var_pull(oracles[f], k.ctypes.data, B, want.ctypes.data)
            worst = max(worst, float(np.abs(want - out[:, sl[f]].cpu().numpy()).max()))
        plan.push_update(ids, grads)
        torch.cuda.synchronize()
        be.engine.check()
        allids = [None] * world
        allg = [None] * world
        dist.all_gather_object(allids, ids.cpu())
        dist.all_gather_object(allg, grads.cpu())
        for f, (dim, vocab, is_hash) in enumerate(specs):
            for r in range(world):
                k = np.ascontiguousarray(allids[r][:, f].numpy(), dtype=np.uint64)
                gg = np.ascontiguousarray(allg[r][:, sl[f]].numpy(), dtype=np.float32)
                lib.exb_var_push(oracles[f], k.ctypes.data, B, gg.ctypes.data, None)
            lib.exb_var_update(oracles[f])
    assert worst < 5e-4, worst
    if rank == 0:
        print("[mp] pull/push_update vs oracle ok, max err %.2e, phases %s" % (worst, be.engine.status()[1]["last_push_update_us"]))
    # ---- P2P all-reduce
    from openembedding_b200.ops.p2p_allreduce import P2PAllReduce
    mk = lambda: torch.randn(1 << 20, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    flat = mk()
    ref = flat.clone()
    dist.all_reduce(ref)
    ar = P2PAllReduce(ctx, flat)
    for _ in range(3):
        flat.copy_(mk())
        ar()
    torch.cuda.synchronize()
    assert ar.status() == 0
    assert torch.allclose(flat, ref, atol=1e-5, rtol=1e-5), float((flat - ref).abs().max())
    if rank == 0:
        print("[mp] p2p all-reduce ok")
    # ---- collective checkpoint, reload
    from openembedding_b200 import checkpoint as ck
    d = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    path = d[0] + "/model"
    ck.save_model(ctx, path, include_optimizer=True)
    probe = torch.stack([torch.arange(B) % min(v, 400) * (1000003 if h else 1) + (7 if h else 0)
                         for (_, v, h) in specs], 1).contiguous().to(dev)
    before = plan.pull(probe).clone()
    plan.push_update(probe, torch.randn(B, plan.io_stride, device=dev))
    torch.cuda.synchronize()
    assert not torch.equal(plan.pull(probe)[:, sl[0]], before[:, sl[0]])
    ck.load_model(ctx, path)
    after = plan.pull(probe)
    torch.cuda.synchronize()
    if True:
        for f, (dim, vocab, is_hash) in enumerate(specs):
            dlt = (after[:, sl[f]] - before[:, sl[f]]).abs().max(dim=1)[0]
            bad = torch.nonzero(dlt > 0).reshape(-1)
            if bad.numel():
                bid = probe[bad, f]
                own = (metas[f].shard_base + bid % metas[f].shard_num) % world
                print("[mp][rank %d] table %d (dim %d hash %s base %d): %d bad rows, owners %s, first ids %s, maxdiff %.3g" % (
                    rank, f, dim, is_hash, metas[f].shard_base, bad.numel(), torch.bincount(own.cpu(), minlength=world).tolist(),
                    bid[:6].tolist(), float(dlt.max())), flush=True)
    for f in range(len(specs)):   # padding columns between features are never written: compare features only
        assert torch.equal(after[:, sl[f]], before[:, sl[f]]), (f, float((after[:, sl[f]] - before[:, sl[f]]).abs().max()))
    be.engine.check()
    dist.barrier()
    if rank == 0:
        print("[mp] checkpoint save/load (world=%d) ok" % world)
        shutil.rmtree(d[0], ignore_errors=True)
        print("MP_GPU_CHECK_PASSED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
