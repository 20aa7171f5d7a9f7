"""Criteo LR with ONE hashed embedding table (counterpart of the reference's
examples/criteo_lr_subclass.py): input_dim=-1 -> ids in [0, 2**63) live in a hash table."""
import argparse
import os
import sys

import pandas
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openembedding_b200.torch as embed  # noqa: E402
from openembedding_b200.models.ctr import CriteoLR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--data", default="")
ap.add_argument("--checkpoint", default="")   # include optimizer
ap.add_argument("--load", default="")
ap.add_argument("--save", default="")         # not include optimizer
ap.add_argument("--epochs", type=int, default=5)
ap.add_argument("--batch_size", type=int, default=8)
args = ap.parse_args()

if args.data:
    data = pandas.read_csv(args.data)
else:
    from make_sample_data import make
    data = make(100)
sparse = torch.stack([torch.tensor((data["C%d" % i].astype("int64") * 1000003 + i * 1000000007).values) for i in range(1, 27)], 1)
dense = torch.tensor(data[["I%d" % i for i in range(1, 14)]].values, dtype=torch.float32)
label = torch.tensor(data["label"].values, dtype=torch.float32)

model = embed.distributed_model(CriteoLR(num_shards=16))
optimizer = embed.distributed_optimizer(torch.optim.Adam(model.parameters()))
if args.load:
    model.load_weights(args.load)
dev = next(model.out.parameters()).device
for epoch in range(args.epochs):
    total = 0.0
    for i in range(0, len(label), args.batch_size):
        sl = slice(i, i + args.batch_size)
        logit = model(sparse[sl], dense[sl])
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, label[sl].to(dev))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        total += float(loss.detach()) * (sl.stop - sl.start)
    print("epoch %d loss %.4f" % (epoch + 1, total / len(label)))
    if args.checkpoint:
        model.save_weights(args.checkpoint + str(epoch + 1))
if args.save:
    model.save(args.save, include_optimizer=False)
