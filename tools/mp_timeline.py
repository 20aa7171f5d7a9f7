"""Per-stage device time of the fused DeepFM step at N ranks (CUDA events, eager launches).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mp_timeline.py [--overlap 0|1]

ncu cannot profile a multi-rank step (kernel replay breaks the cross-GPU barriers), so this is the
multi-GPU counterpart of tools/step_timeline.py. Prints mean microseconds per stage on every rank.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--overlap", default="0")
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--batch", type=int, default=4096)
a = ap.parse_args()
os.environ["EXB_OVERLAP"] = a.overlap
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
from openembedding_b200.context import get_context  # noqa: E402
from openembedding_b200.models.ctr import CRITEO_1TB_VOCAB_20M  # noqa: E402
from openembedding_b200.models.fused_dense import FusedCTR  # noqa: E402

ctx = get_context()
dev = ctx.device
m = FusedCTR(CRITEO_1TB_VOCAB_20M, embedding_dim=64, model="deepfm", batch=a.batch, cache_threshold=a.batch)
g = torch.Generator().manual_seed(1 + ctx.rank)
batches = []
for _ in range(8):
    u = torch.rand(a.batch, 26, generator=g)
    ids = torch.stack([(torch.exp(u[:, f] * torch.log(torch.tensor(float(v)))) - 1).long().clamp_(0, v - 1)
                       for f, v in enumerate(CRITEO_1TB_VOCAB_20M)], 1).contiguous().to(dev)
    batches.append((ids, torch.rand(a.batch, 13, generator=g).to(dev), (torch.rand(a.batch, generator=g) < 0.3).float().to(dev)))
for i in range(10):
    m.forward_backward(*batches[i % 8])
torch.cuda.synchronize()
acc, tot = {}, 0.0
for i in range(a.steps):
    if world > 1:
        dist.barrier()
    m._trace = []
    m.forward_backward(*batches[i % 8])
    torch.cuda.synchronize()
    tr = m._trace
    for (n0, e0), (n1, e1) in zip(tr[:-1], tr[1:]):
        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) * 1e3
    tot += tr[0][1].elapsed_time(tr[-1][1]) * 1e3
m._trace = None
line = "rank %d overlap=%s  total %.1f us | " % (ctx.rank, a.overlap, tot / a.steps)
line += "  ".join("%s %.1f" % (k, v / a.steps) for k, v in acc.items())
st = ctx.backend.engine.status()[1]
print(line, flush=True)
if ctx.rank == 0 and st:
    print("push phases (last step, us):", st.get("last_push_update_us"), flush=True)
    if m._ar is not None:
        print("all-reduce phases (last step, us):", m._ar.phases_us(), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
