"""torchrun script (gloo, CPU): the host-DRAM tier on a multi-rank job gives bit-identical results to the untiered
table (every rank promotes / evicts / writes back only the rows it owns)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(tier_rows):
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context, reset_context
    reset_context()
    ctx = get_context()
    rank, world = ctx.rank, ctx.world
    emb = embed.Embedding(-1, 6, embeddings_initializer={"category": "uniform", "minval": -0.5, "maxval": 0.5},
                          host_tier_rows=tier_rows)
    opt = embed.distributed_optimizer(torch.optim.Adagrad(emb.parameters(), lr=0.1, initial_accumulator_value=0.1))
    g = torch.Generator().manual_seed(3)
    for step in range(25):
        ids = torch.randint(0, 400, (16 * world,), generator=g) * 7 + 1
        mine = ids[rank * 16:(rank + 1) * 16]
        loss = (emb(mine) ** 2).sum()
        opt.zero_grad()
        loss.backward()
        opt.step()
    rows = emb(torch.arange(0, 400) * 7 + 1).detach().clone()
    stats = dict(emb.variable.tier.stats) if tier_rows else {}
    reset_context()
    return rows, stats


def main():
    dist.init_process_group("gloo")
    import openembedding_b200 as oe
    oe.flags.device = "cpu"
    plain, _ = run(None)
    tiered, stats = run(48)
    assert torch.equal(plain, tiered), (plain - tiered).abs().max()
    assert stats["evictions"] > 0 and stats["writebacks"] > 0, stats
    both = [torch.zeros_like(tiered) for _ in range(dist.get_world_size())]
    dist.all_gather(both, tiered)
    assert all(torch.equal(both[0], b) for b in both)
    if dist.get_rank() == 0:
        print("MP_CPU_TIER_CHECK_PASSED", stats)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
