"""Python face of the hand-written tcgen05 GEMM (``csrc/cuda/gemm_tcgen05.cu``).

``gemm_nt(A, B, ...)`` computes ``A[M,K] @ B[N,K].T`` on bf16 operands with one of the
fused epilogues. Operands must be K-padded to a multiple of 64 and have leading
dimensions that are multiples of 8 elements (the dense engine allocates them that way).
"""
import ctypes
from ctypes import c_int, c_longlong, c_uint64

import torch

from .. import _native

EPI_FWD, EPI_DX, EPI_DW, EPI_DX_FM = 0, 1, 2, 3
_proto_done = False


def _lib():
    global _proto_done
    lib = _native.cuda()
    if not _proto_done:
        lib.exb_gemm_bf16_nt.restype = c_int
        lib.exb_gemm_bf16_nt.argtypes = [c_uint64, c_longlong, c_uint64, c_longlong, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_uint64, c_longlong, c_uint64, c_longlong, c_uint64, c_longlong,
                                         c_uint64, c_uint64, c_uint64, c_longlong, c_int, c_int, c_int, c_uint64,
                                         c_uint64]
        lib.exb_gemm_bf16_tn.restype = c_int
        lib.exb_gemm_bf16_tn.argtypes = [c_uint64, c_longlong, c_uint64, c_longlong, c_int, c_int, c_int, c_uint64,
                                         c_longlong, c_int, c_uint64]
        lib.exb_gemm_last_error.restype = ctypes.c_char_p
        _proto_done = True
    return lib


def _p(t):
    return t.data_ptr() if t is not None else 0


def gemm_nt(A, B, M, N, K, out, mode=EPI_FWD, relu=False, ones_col=-1, outT=None, mask=None, dlogit=None, S=None,
            emb=None, fm_cols=0, D=1, splits=1, stream=None, dbg=None):
    """A: [>=M, lda] bf16 view, B: [>=N, ldb] bf16 view (both row-major, K contiguous)."""
    lib = _lib()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    st = stream if stream is not None else torch.cuda.current_stream(A.device).cuda_stream
    rc = lib.exb_gemm_bf16_nt(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, mode, int(relu), ones_col,
                              out.data_ptr(), out.stride(0), _p(outT), outT.stride(0) if outT is not None else 0,
                              _p(mask), mask.stride(0) if mask is not None else 0, _p(dlogit), _p(S), _p(emb),
                              emb.stride(0) if emb is not None else 0, fm_cols, D, splits, st, _p(dbg))
    if rc != 0:
        raise RuntimeError("exb_gemm_bf16_nt: " + lib.exb_gemm_last_error().decode())
    return out


def gemm_tn(A, B, M, N, K, out, splits=1, stream=None):
    """out[M, N] (fp32) += A[K, M].T @ B[K, N] with A, B row-major bf16 (K rows): the weight-gradient
    product straight from batch-major activations -- both operands are fed to the tensor core as
    MN-major tiles, so no transposed copies have to be materialised. ``out`` must be zeroed (or
    hold the value to accumulate onto); split-K partial sums arrive by TMA reduce-add."""
    lib = _lib()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    st = stream if stream is not None else torch.cuda.current_stream(A.device).cuda_stream
    rc = lib.exb_gemm_bf16_tn(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, out.data_ptr(),
                              out.stride(0), splits, st)
    if rc != 0:
        raise RuntimeError("exb_gemm_bf16_tn: " + lib.exb_gemm_last_error().decode())
    return out
