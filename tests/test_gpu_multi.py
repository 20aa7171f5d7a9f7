"""Real multi-GPU correctness (CUDA IPC over NVLink, one process per GPU), collected by ``pytest -m gpu``.

Each test self-launches ``torchrun`` on this box when at least two GPUs are visible and is skipped otherwise
(the single-GPU coverage of the multi-rank protocol is the virtual-rank suite in test_gpu_sparse_engine.py).
Reference model: openembedding/entry/c_api_test.cpp runs every case on 1..9 forked nodes."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script, *args, timeout=600, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, script)] + list(args)
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=e)
    return p.returncode, p.stdout


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on the box")
@pytest.mark.parametrize("n", [2, 4, 8])
def test_mp_engine_allreduce_checkpoint(n):
    """fused pull / push+update vs the CPU oracle, P2P all-reduce vs NCCL, collective checkpoint + re-shard load"""
    if _ngpu() < n:
        pytest.skip("box has %d GPUs" % _ngpu())
    rc, out = _torchrun(n, "tests/mp_gpu_check.py")
    assert rc == 0 and "MP_GPU_CHECK_PASSED" in out, out[-4000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on the box")
def test_mp_fused_step_two_ranks():
    """the product step (FusedCTR + CUDA graph) trains on 2 ranks and the ranks stay bit-identical replicas"""
    rc, out = _torchrun(2, "tests/mp_gpu_fused_check.py")
    assert rc == 0 and "MP_GPU_FUSED_PASSED" in out and "rider 1" in out, out[-4000:]
    # the dense all-reduce riding on the push kernel gives the parameters of the stand-alone kernel (not bit for bit:
    # the sparse gradient accumulation uses float atomics, whose order varies from run to run)
    rc2, out2 = _torchrun(2, "tests/mp_gpu_fused_check.py", env={"EXB_AR_RIDER": "0"})
    assert rc2 == 0 and "MP_GPU_FUSED_PASSED" in out2 and "rider 0" in out2, out2[-4000:]
    tsum = lambda o: float(o.split("theta_sum ")[1].split()[0])
    assert abs(tsum(out) - tsum(out2)) < 1e-6 * abs(tsum(out2)), (tsum(out), tsum(out2))


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on the box")
@pytest.mark.parametrize("v2", ["0", "1"])
def test_mp_fused_step_prefetch(v2):
    """same step prefetching the next batch in its tail, with the v1 / planned (v2) sparse kernels at world 2"""
    rc, out = _torchrun(2, "tests/mp_gpu_fused_check.py", env={"EXB_SPARSE_V2": v2, "EXB_TEST_PREFETCH": "1"})
    assert rc == 0 and "MP_GPU_FUSED_PASSED" in out, out[-4000:]
