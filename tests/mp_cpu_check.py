"""BASELINE config 1 (plumbing): Criteo LR (examples/criteo_lr_subclass) with a 1M-row table,
world_size=2, CPU + gloo. Also: sum-gradient equivalence with a single process and a
checkpoint written by 2 ranks."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import openembedding_b200 as oe
    oe.flags.device = "cpu"
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context
    from openembedding_b200.models.ctr import CriteoLR
    ctx = get_context()
    assert ctx.world == world and ctx.backend.name == "cpu"
    torch.manual_seed(0)
    model = CriteoLR(num_sparse=26, num_dense=13, input_dim=1000000, num_shards=16)
    for p in model.out.parameters():           # BroadcastGlobalVariablesCallback(0)
        dist.broadcast(p.data, src=0)
    opt = embed.distributed_optimizer(torch.optim.Adam(model.parameters(), lr=0.1))
    g = torch.Generator().manual_seed(7)       # same stream on all ranks, each takes its slice
    losses = []
    for step in range(40):
        ids = torch.randint(0, 1000000, (32 * world, 26), generator=g)
        ids[:, 0] = ids[:, 0] % 50              # a hot column: duplicate ids within and across ranks
        dense = torch.rand(32 * world, 13, generator=g)
        y = (ids[:, 0] % 2).float()
        sl = slice(rank * 32, (rank + 1) * 32)
        logit = model(ids[sl], dense[sl])
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, y[sl])
        opt.zero_grad()
        loss.backward()
        for p in model.out.parameters():       # hvd.DistributedOptimizer(op=Sum) on the dense part
            dist.all_reduce(p.grad)
        opt.step()
        t = loss.detach().clone()
        dist.all_reduce(t)
        losses.append(float(t) / world)
    assert losses[-1] < losses[0] - 0.05, losses
    # pulls agree on every rank (rows live on different ranks)
    probe = torch.arange(0, 50)
    rows = model.embeddings(probe).detach()
    both = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(both, rows)
    assert all(torch.equal(both[0], b) for b in both)
    d = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    embed.save_server_model(model, d[0] + "/ck")
    dist.barrier()
    if rank == 0:
        files = sorted(os.listdir(d[0] + "/ck/0"))
        assert files == ["model_%d_0" % r for r in range(world)], files
        torch.save({"rows": rows, "dir": d[0]}, os.environ.get("EXB_MP_OUT", d[0] + "/out.pt"))
        print("MP_CPU_CHECK_PASSED", losses[0], losses[-1])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
