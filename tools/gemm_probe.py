import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openembedding_b200.ops.gemm import EPI_DX_FM, EPI_FWD, EPI_DW, gemm_nt
dev = "cuda"
def mk(r, c): return (torch.randn(r, c, device=dev) * 0.1).to(torch.bfloat16)
B = 4096
cases = {
  "fwd1  M4096 N448 K1728": dict(A=mk(B, 1728), Bm=mk(448, 1728), M=B, N=448, K=1728, mode=EPI_FWD, out=torch.zeros(B, 448, device=dev, dtype=torch.bfloat16), outT=torch.zeros(448, B, device=dev, dtype=torch.bfloat16)),
  "fwd1-noT M4096 N448 K1728": dict(A=mk(B, 1728), Bm=mk(448, 1728), M=B, N=448, K=1728, mode=EPI_FWD, out=torch.zeros(B, 448, device=dev, dtype=torch.bfloat16)),
  "fwd1-nostore M4096 N448 K1728": dict(A=mk(B, 1728), Bm=mk(448, 1728), M=B, N=448, K=1728, mode=EPI_FWD, out=torch.zeros(B, 448, device=dev, dtype=torch.bfloat16), outT=torch.zeros(448, B, device=dev, dtype=torch.bfloat16), nostore=True),
  "dX1   M4096 N1728 K448": dict(A=mk(B, 448), Bm=mk(1728, 448), M=B, N=1728, K=448, mode=EPI_DX_FM, out=torch.zeros(B, 1756, device=dev), fm=True),
  "dX1nf M4096 N1728 K448": dict(A=mk(B, 448), Bm=mk(1728, 448), M=B, N=1728, K=448, mode=EPI_DX_FM, out=torch.zeros(B, 1756, device=dev), fm=False),
  "fwd2  M4096 N448 K448": dict(A=mk(B, 448), Bm=mk(448, 448), M=B, N=448, K=448, mode=EPI_FWD, out=torch.zeros(B, 448, device=dev, dtype=torch.bfloat16)),
  "dW1   M448 N1728 K4096": dict(A=mk(448, B), Bm=mk(1728, B), M=448, N=1728, K=B, mode=EPI_DW, out=torch.zeros(448, 1728, device=dev), splits=8),
}
emb = torch.randn(B, 1756, device=dev); S = torch.randn(B, 64, device=dev); dl = torch.randn(B, device=dev)
for name, c in cases.items():
    dbg = torch.zeros(8, dtype=torch.int64, device=dev)
    kw = {}
    if c.get("nostore"):
        kw = dict(fm_cols=-7)
    if c["mode"] == EPI_DX_FM:
        kw = dict(dlogit=dl, S=S, emb=emb, fm_cols=1664 if c.get("fm") else 0, D=64)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gemm_nt(c["A"], c["Bm"], c["M"], c["N"], c["K"], c["out"], mode=c["mode"], relu=True, outT=c.get("outT"), splits=c.get("splits", 1), dbg=dbg, **kw)
        e1.record(); torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    print("%s: total %.1f us | CTA0: setup %.2f, first-load %.2f, mainloop %.2f, epilogue %.2f, teardown %.2f (us)" % (
        name, e0.elapsed_time(e1) * 1e3, (t[1]-t[0])/1e3, (t[2]-t[1])/1e3, (t[3]-t[2])/1e3, (t[4]-t[3])/1e3, (t[5]-t[4])/1e3))
