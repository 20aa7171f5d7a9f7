import ctypes, sys, torch
sys.path.insert(0, '.')
from openembedding_b200.ops.sparse_engine import CudaEngine
from openembedding_b200 import _native
dev = torch.device('cuda', 0)
e = CudaEngine(0, 0, 1)
t = e.add_table(8, 0, True, capacity=1024)
e.set_initializer(t, {"category": "uniform", "minval": -0.5, "maxval": 0.5}, 0)
e.set_optimizer(t, {"category": "adagrad", "learning_rate": 0.1})
e.alloc(t); e.commit()
lib = e.lib
h = lib.exb_tier_create(e.h, t, 1 << 14)
assert h
st = torch.cuda.current_stream().cuda_stream
A = (torch.arange(600, dtype=torch.int64) * 7919 + 3).to(dev)
lib.exb_tier_admit(h, A.data_ptr(), A.numel(), 1, st)
torch.cuda.synchronize(); e.check()
w, s = e.gather_rows(t, A)
print('init w', w[:2], 's', s[:2], 'size', e.table_size(t))
w2, s2 = w * 2 + 1, s + 5
e.scatter_rows(t, A, w2, s2)
wq, sq = e.gather_rows(t, A)
print('scatter ok', torch.equal(wq, w2), torch.equal(sq, s2))
out = (ctypes.c_uint64 * 8)()
lib.exb_tier_evict(h, 100, 5, st); torch.cuda.synchronize(); e.check()
lib.exb_tier_stats(h, out); print('stats after evict', list(out), 'size', e.table_size(t))
lib.exb_tier_admit(h, A.data_ptr(), A.numel(), 6, st); torch.cuda.synchronize(); e.check()
lib.exb_tier_stats(h, out); print('stats after readmit', list(out), 'size', e.table_size(t))
wr, sr = e.gather_rows(t, A)
print('roundtrip w', torch.equal(wr, w2), 's', torch.equal(sr, s2))
if not torch.equal(wr, w2):
    bad = (wr != w2).any(1).nonzero().reshape(-1)
    print('bad rows', bad.numel(), bad[:10].tolist(), wr[bad[:2]], w2[bad[:2]])
# partial eviction: touch half at work 7, evict to 350 at work 8
B = A[:300].contiguous()
lib.exb_tier_admit(h, B.data_ptr(), B.numel(), 7, st)
lib.exb_tier_evict(h, 350, 8, st); torch.cuda.synchronize(); e.check()
lib.exb_tier_stats(h, out); print('stats after partial evict', list(out), 'size', e.table_size(t))
wr, sr = e.gather_rows(t, B)
print('survivors intact', torch.equal(wr, w2[:300]), torch.equal(sr, s2[:300]))
lib.exb_tier_admit(h, A.data_ptr(), A.numel(), 9, st); torch.cuda.synchronize(); e.check()
wr, sr = e.gather_rows(t, A)
print('all back', torch.equal(wr, w2), torch.equal(sr, s2))
