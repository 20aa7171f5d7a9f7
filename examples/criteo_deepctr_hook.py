"""`distributed_model` on a plain PyTorch network -- counterpart of the reference's
examples/criteo_deepctr_hook.py (there: swap deepctr's Embedding for the PS embedding).
Every nn.Embedding becomes a server-side table (small ones stay replicated), the wrapped
optimizer drives the sparse update, and the model gains save / load_weights /
save_as_original_model."""
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openembedding_b200.torch as embed  # noqa: E402
from make_sample_data import make  # noqa: E402


class Net(nn.Module):
    def __init__(self, vocab, dim=8):
        super().__init__()
        self.embs = nn.ModuleList([nn.Embedding(v, dim) for v in vocab])
        self.mlp = nn.Sequential(nn.Linear(len(vocab) * dim + 13, 64), nn.ReLU(), nn.Linear(64, 1))

    def forward(self, ids, dense):
        e = [emb(ids[:, i]) for i, emb in enumerate(self.embs)]
        return self.mlp(torch.cat(e + [dense.to(e[0].device)], 1)).squeeze(-1)


data = make(128)
vocab = [int(data["C%d" % i].max()) + 1 for i in range(1, 27)]
ids = torch.tensor(data[["C%d" % i for i in range(1, 27)]].values)
dense = torch.tensor(data[["I%d" % i for i in range(1, 14)]].values, dtype=torch.float32)
label = torch.tensor(data["label"].values, dtype=torch.float32)

model = embed.distributed_model(Net(vocab), sparse_as_dense_size=64)
opt = embed.distributed_optimizer(torch.optim.Adagrad(model.parameters(), lr=0.05, initial_accumulator_value=0.1))
for epoch in range(3):
    for i in range(0, 128, 32):
        sl = slice(i, i + 32)
        logit = model(ids[sl], dense[sl])
        loss = nn.functional.binary_cross_entropy_with_logits(logit, label[sl].to(logit.device))
        opt.zero_grad(); loss.backward(); opt.step()
    print("epoch", epoch + 1, "loss %.4f" % float(loss))
out = os.environ.get("EXB_EXAMPLE_OUT", "/tmp/exb_hook_example")
model.save(out, include_optimizer=False)
model.save_as_original_model(out + "/standalone.pt")
print("saved to", out)
