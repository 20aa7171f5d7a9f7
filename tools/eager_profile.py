"""torch.profiler kernel table of a few eager-zoo steps (xdeepfm / dcn): which kernels make the step
usage: python tools/eager_profile.py --model xdeepfm --dim 9"""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openembedding_b200 as oe
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="xdeepfm"); ap.add_argument("--dim", type=int, default=9)
ap.add_argument("--batch", type=int, default=4096); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
oe.flags.device = "cuda"
from openembedding_b200.context import get_context
from openembedding_b200.models.ctr import CRITEO_1TB_VOCAB_20M, CTRModel
from openembedding_b200.models.trainer import Trainer
ctx = get_context()
vocab = CRITEO_1TB_VOCAB_20M
model = CTRModel(vocab, num_dense=13, embedding_dim=a.dim, model=a.model, batch=a.batch,
                 sparse_optimizer={"category": "adagrad"}, cache_threshold=4096)
tr = Trainer(model, use_graph=False, dense_optimizer={"category": "adagrad"})
host = B.make_batches(torch, vocab, 13, a.batch, 4, 1.0, 1000, ctx.device)
devb = [(i.to(ctx.device), d.to(ctx.device), l.to(ctx.device)) for i, d, l in host]
for s in range(5):
    tr.step(*devb[s % 4])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(a.steps):
        tr.step(*devb[s % 4])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
