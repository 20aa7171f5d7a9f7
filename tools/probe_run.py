import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openembedding_b200 as oe
from openembedding_b200.context import get_context
from openembedding_b200.models.ctr import CRITEO_KAGGLE_VOCAB, CTRModel
oe.flags.device = "cuda"
ctx = get_context()
vocab = CRITEO_KAGGLE_VOCAB
m = CTRModel(vocab, embedding_dim=64, model="deepfm", batch=4096, cache_threshold=4096)
g = m.sparse.group
dev = ctx.device
gen = torch.Generator().manual_seed(0)
v = torch.tensor(vocab, dtype=torch.float64)
for it in range(6):
    u = torch.rand((4096, 26), generator=gen, dtype=torch.float64)
    ids = (torch.floor(torch.exp(u * torch.log(v))) - 1).clamp_(min=0).to(torch.int64)
    ids = ((ids * 2654435761 + 12345) % torch.tensor(vocab)).to(dev)
    grads = torch.randn(4096, g.io_stride, device=dev)
    torch.cuda.synchronize()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record(); out = g.pull(ids); e1.record(); g.push_update(ids, grads); e2.record()
    torch.cuda.synchronize()
    st = ctx.backend.engine.status()[1]
    p = st["probe"]
    print("iter", it, "pull %.1f us push %.1f us" % (e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3), st["last_push_update_us"])
    print("   probe(us): stage->task %.2f  cas %.2f  cnt %.2f  move %.2f  grid %s" % (
        (p[0] - p[5]) / 1e3, (p[2] - p[0]) / 1e3, (p[3] - p[2]) / 1e3, (p[4] - p[3]) / 1e3, g.grid()))
    print("   probe5(us): resolve %.2f apply %.2f" % ((p[7] - p[6]) / 1e3, (p[8] - p[7]) / 1e3))
