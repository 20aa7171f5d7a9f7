"""tcgen05 GEMM vs a plain PyTorch fp32 reference of the same op, every fused epilogue."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, ld=None, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ld = ld or cols
    buf = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)
    buf[:, :cols] = (torch.randn(rows, cols, device="cuda", generator=g) * scale).to(torch.bfloat16)
    return buf


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 192), (4096, 448, 1728), (4096, 448, 448), (300, 100, 128)])
def test_fwd_relu_ones_transposed(M, N, K):
    from openembedding_b200.ops.gemm import EPI_FWD, gemm_nt
    Np = (N + 63) // 64 * 64
    A, B = _mk(M, K, seed=1), _mk(Np, K, scale=0.1, seed=2)
    B[N:] = 0
    out = torch.full((M, Np), 7.0, device="cuda", dtype=torch.bfloat16)
    Mp = (M + 7) // 8 * 8        # TMA store: leading dimensions must be multiples of 16 bytes
    outT = torch.full((Np, Mp), 7.0, device="cuda", dtype=torch.bfloat16)[:, :M]
    ones_col = N - 1
    gemm_nt(A, B, M, N, K, out, mode=EPI_FWD, relu=True, ones_col=ones_col, outT=outT)
    torch.cuda.synchronize()
    ref = torch.relu(A.float() @ B[:N].float().t())
    ref[:, ones_col] = 1.0
    assert torch.allclose(out[:, :N].float(), ref, atol=2e-2, rtol=2e-2), (out[:, :N].float() - ref).abs().max()
    assert torch.equal(outT[:N].t().contiguous(), out[:, :N].contiguous())


def test_dx_mask():
    from openembedding_b200.ops.gemm import EPI_DX, gemm_nt
    M, N, K = 512, 448, 448
    dZ, WT = _mk(M, K, seed=3), _mk(N, K, scale=0.1, seed=4)
    H = _mk(M, N, seed=5)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    outT = torch.zeros(N, M, device="cuda", dtype=torch.bfloat16)
    gemm_nt(dZ, WT, M, N, K, out, mode=EPI_DX, ones_col=N - 1, outT=outT, mask=H)
    torch.cuda.synchronize()
    ref = (dZ.float() @ WT.float().t()) * (H.float() > 0)
    ref[:, N - 1] = 0
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=3e-2)
    assert torch.equal(outT.t().contiguous(), out)


@pytest.mark.parametrize("splits", [1, 4, 16])
def test_dw_splitk(splits):
    from openembedding_b200.ops.gemm import EPI_DW, gemm_nt
    M, N, K = 448, 1728, 4096      # dW1 = dZ1^T[448,B] @ A0^T[1728,B]^T
    A, B = _mk(M, K, scale=0.1, seed=6), _mk(N, K, scale=0.1, seed=7)
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    gemm_nt(A, B, M, N, K, out, mode=EPI_DW, splits=splits)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert torch.allclose(out, ref, atol=5e-2, rtol=2e-2), (out - ref).abs().max()


def test_dx_fm():
    from openembedding_b200.ops.gemm import EPI_DX_FM, gemm_nt
    Bsz, N, K, D, F = 256, 1728, 448, 64, 26
    dZ, WT = _mk(Bsz, K, seed=8), _mk(N, K, scale=0.1, seed=9)
    g = torch.Generator(device="cuda").manual_seed(10)
    emb = torch.randn(Bsz, 1800, device="cuda", generator=g)
    S = emb[:, :F * D].reshape(Bsz, F, D).sum(1).contiguous()
    dl = torch.randn(Bsz, device="cuda", generator=g)
    out = torch.zeros(Bsz, 1800, device="cuda")
    gemm_nt(dZ, WT, Bsz, N, K, out, mode=EPI_DX_FM, dlogit=dl, S=S, emb=emb, fm_cols=F * D, D=D)
    torch.cuda.synchronize()
    ref = dZ.float() @ WT.float().t()
    fm = dl[:, None, None] * (S[:, None, :] - emb[:, :F * D].reshape(Bsz, F, D))
    ref[:, :F * D] += fm.reshape(Bsz, -1)
    assert torch.allclose(out[:, :N], ref, atol=5e-2, rtol=2e-2), (out[:, :N] - ref).abs().max()
    assert float(out[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K,splits", [(128, 64, 64, 1), (64, 64, 128, 1), (448, 1728, 4096, 8), (448, 448, 4096, 8),
                                          (100, 72, 256, 2)])
def test_dw_mn_major(M, N, K, splits):
    """out = A[K,M]^T @ B[K,N] from batch-major operands (MN-major UMMA tiles, no transposed copies)"""
    from openembedding_b200.ops.gemm import gemm_tn
    ldm, ldn = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A, B = _mk(K, M, ld=ldm, scale=0.1, seed=11), _mk(K, N, ld=ldn, scale=0.1, seed=12)
    out = torch.zeros(M, ldn, device="cuda", dtype=torch.float32)
    gemm_tn(A[:, :M], B[:, :N], M, N, K, out[:, :N] if ldn == N else out, splits=splits)
    torch.cuda.synchronize()
    ref = A[:, :M].float().t() @ B[:, :N].float()
    assert torch.allclose(out[:, :N], ref, atol=5e-2, rtol=2e-2), (out[:, :N] - ref).abs().max()
