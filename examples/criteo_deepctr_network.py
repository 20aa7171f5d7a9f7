"""DeepFM / WDL / xDeepFM / DCN on Criteo-shaped data -- counterpart of the reference's
examples/criteo_deepctr_network{,_mirrored,_mpi}.py and test/benchmark/criteo_deepctr.py.

single GPU / CPU :  python examples/criteo_deepctr_network.py --model DeepFM
multi GPU        :  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
                        examples/criteo_deepctr_network.py --model DeepFM --fused
(torchrun replaces horovodrun / MirroredStrategy / mpirun of the reference: one rank per GPU,
dense gradients are summed across ranks, the embedding tables are row-sharded over the GPUs)
"""
import argparse
import os
import sys

import pandas
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openembedding_b200 as oe  # noqa: E402
import openembedding_b200.torch as embed  # noqa: E402
from openembedding_b200.context import get_context  # noqa: E402
from openembedding_b200.models.ctr import CTRModel  # noqa: E402
from openembedding_b200.models.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--data", default="")
ap.add_argument("--model", default="DeepFM", choices=["LR", "WDL", "DeepFM", "xDeepFM", "DCN"])
ap.add_argument("--optimizer", default="Adagrad", choices=["Adam", "Adagrad", "Ftrl", "SGD"])
ap.add_argument("--embedding_dim", type=int, default=9)
ap.add_argument("--batch_size", type=int, default=16)
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--cache", action="store_true", help="replicate tables smaller than the batch (sparse_as_dense)")
ap.add_argument("--fused", action="store_true", help="whole step on the hand-written kernels (CUDA, DeepFM/WDL)")
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--checkpoint", default="")
ap.add_argument("--load", default="")
ap.add_argument("--save", default="")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and not args.cpu
    if use_cuda:
        torch.cuda.set_device(local)
    dist.init_process_group("nccl" if use_cuda else "gloo")
oe.flags.device = "cpu" if args.cpu else "auto"
ctx = get_context()

if args.data:
    data = pandas.read_csv(args.data)
else:
    from make_sample_data import make
    data = make(128)
vocab = [int(data["C%d" % i].max()) + 1 for i in range(1, 27)]
n = len(data) // (world * args.batch_size) * args.batch_size
part = data.iloc[ctx.rank * n:(ctx.rank + 1) * n]
ids = torch.tensor(part[["C%d" % i for i in range(1, 27)]].values, dtype=torch.int64)
dense = torch.tensor(part[["I%d" % i for i in range(1, 14)]].values, dtype=torch.float32)
label = torch.tensor(part["label"].values, dtype=torch.float32)

sparse_opt = {"category": args.optimizer.lower()}
cache = args.batch_size if args.cache else 0
if args.fused:
    from openembedding_b200.models.fused_dense import FusedCTR, FusedTrainer
    model = FusedCTR(vocab, embedding_dim=args.embedding_dim, model=args.model.lower(), batch=args.batch_size,
                     sparse_optimizer=sparse_opt, cache_threshold=cache)
    trainer = FusedTrainer(model, use_graph=True)
else:
    model = CTRModel(vocab, embedding_dim=args.embedding_dim, model=args.model.lower(), batch=args.batch_size,
                     sparse_optimizer=sparse_opt, cache_threshold=cache,
                     compute_dtype=torch.float32 if ctx.device.type == "cpu" else torch.bfloat16)
    trainer = Trainer(model, use_graph=False)
if args.load:
    embed.load_server_model(model, args.load)
for epoch in range(args.epochs):
    tot, cnt = 0.0, 0
    for i in range(0, n, args.batch_size):
        sl = slice(i, i + args.batch_size)
        loss = trainer.step(ids[sl].contiguous().to(ctx.device), dense[sl].to(ctx.device), label[sl].to(ctx.device))
        tot += float(loss)
        cnt += 1
    if ctx.rank == 0:
        print("epoch %d loss %.4f" % (epoch + 1, tot / max(cnt, 1)))
    if args.checkpoint:
        embed.save_server_model(model, args.checkpoint + str(epoch + 1))          # include optimizer
if args.save:
    embed.save_server_model(model, args.save, include_optimizer=False)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
