"""torchrun script (gloo, CPU, world=2): the model-level API on a multi-rank job --
distributed_model + distributed_optimizer, model.save / load_weights, save_as_original_model
(standalone export written by rank 0 from rows pulled across ranks), host-tier persist / restore."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.big = torch.nn.Embedding(3000, 6)
        self.small = torch.nn.Embedding(12, 3)
        self.out = torch.nn.Linear(9, 1)

    def forward(self, a, b):
        return self.out(torch.cat([self.big(a), self.small(b)], -1)).squeeze(-1)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import openembedding_b200 as oe
    oe.flags.device = "cpu"
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context, reset_context
    torch.manual_seed(0)
    model = embed.distributed_model(Net(), sparse_as_dense_size=64)
    for p in model.parameters():
        dist.broadcast(p.data, src=0)
    opt = embed.distributed_optimizer(torch.optim.Adagrad(model.parameters(), lr=0.1, initial_accumulator_value=0.1))
    g = torch.Generator().manual_seed(5)

    def step():
        a = torch.randint(0, 3000, (8 * world,), generator=g)
        b = torch.randint(0, 12, (8 * world,), generator=g)
        y = (a % 2).float()
        sl = slice(rank * 8, (rank + 1) * 8)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(model(a[sl], b[sl]), y[sl])
        opt.zero_grad()
        loss.backward()
        for name, p in model.named_parameters():
            if p.grad is not None and not name.startswith("big."):   # dense + replicated small table: hvd Sum
                dist.all_reduce(p.grad)
        opt.step()
    for _ in range(15):
        step()
    d = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    d = d[0]
    probe_a, probe_b = torch.arange(0, 3000, 37), torch.arange(0, 3000, 37) % 12
    want = model(probe_a, probe_b).detach().clone()
    # ---- full checkpoint, keep training, come back
    model.save(d + "/full", include_optimizer=True)
    for _ in range(3):
        step()
    assert not torch.allclose(model(probe_a, probe_b).detach(), want)
    model.load_weights(d + "/full")
    assert torch.allclose(model(probe_a, probe_b).detach(), want, atol=1e-6)
    # ---- standalone export: a plain torch model, loadable without the framework's tables
    model.save_as_original_model(d + "/standalone.pt")
    dist.barrier()
    if rank == 0:
        plain = torch.load(d + "/standalone.pt", weights_only=False)
        assert isinstance(plain.big, torch.nn.Embedding) and plain.big.weight.shape == (3000, 6)
        assert torch.allclose(plain(probe_a, probe_b).detach(), want, atol=1e-6)
    dist.barrier()
    if rank == 0:
        print("MP_CPU_API_CHECK_PASSED")
    reset_context()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
