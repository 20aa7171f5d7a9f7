"""torchrun script (gloo, CPU, world=2): tables with fewer shards than ranks (num_shards=1: array and hash, placed
round-robin on different ranks) next to a fully sharded one; pulls agree on all ranks; checkpoint round trip."""
import os, sys, tempfile
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
import openembedding_b200 as oe
oe.flags.device = "cpu"
import openembedding_b200.torch as embed
from openembedding_b200.context import get_context
e1 = embed.Embedding(500, 4, num_shards=1, embeddings_initializer={"category": "uniform", "minval": -1, "maxval": 1})
e2 = embed.Embedding(-1, 4, num_shards=1, embeddings_initializer={"category": "uniform", "minval": -1, "maxval": 1})
e3 = embed.Embedding(700, 4, embeddings_initializer={"category": "uniform", "minval": -1, "maxval": 1})
params = list(e1.parameters()) + list(e2.parameters()) + list(e3.parameters())
opt = embed.distributed_optimizer(torch.optim.SGD(params, lr=0.5))
g = torch.Generator().manual_seed(1)
for step in range(5):
    ids = torch.randint(0, 500, (8 * world,), generator=g)[rank * 8:(rank + 1) * 8]
    loss = (e1(ids) ** 2).sum() + (e2(ids * 1000003) ** 2).sum() + (e3(ids) ** 2).sum()
    opt.zero_grad(); loss.backward(); opt.step()
probe = torch.arange(0, 500)
rows = torch.cat([e1(probe), e2(probe * 1000003), e3(probe)], 1).detach()
both = [torch.zeros_like(rows) for _ in range(world)]
dist.all_gather(both, rows)
assert all(torch.equal(both[0], b) for b in both)
ctx = get_context()
print(rank, "storages", [(s.storage_id, s.shard_num, s.shard_base) for s in ctx.storages])
d = [tempfile.mkdtemp() if rank == 0 else None]; dist.broadcast_object_list(d, src=0)
class M(torch.nn.Module):
    def __init__(s): super().__init__(); s.e1, s.e2, s.e3 = e1, e2, e3
m = M()
embed.save_server_model(m, d[0] + "/ck")
for step in range(2):
    ids = torch.arange(0, 8) + rank * 8
    loss = (e1(ids) ** 2).sum() + (e2(ids * 1000003) ** 2).sum() + (e3(ids) ** 2).sum()
    opt.zero_grad(); loss.backward(); opt.step()
embed.load_server_model(m, d[0] + "/ck")
rows2 = torch.cat([e1(probe), e2(probe * 1000003), e3(probe)], 1).detach()
assert torch.equal(rows, rows2), (rows - rows2).abs().max()
if rank == 0: print("SHARD1_OK", sorted(os.listdir(d[0] + "/ck")))
dist.destroy_process_group()
