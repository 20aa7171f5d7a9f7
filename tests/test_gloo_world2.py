"""Multi-process CPU path: 2 ranks over gloo (BASELINE.json config 1), then load the 2-rank
checkpoint in a single process (re-shard). Reference analogue: build.sh unit_test
`horovodrun -np 2 ...` + load with a different worker count (build.sh:136-144)."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch


@pytest.mark.parametrize("world", [2, 3])
def test_criteo_lr_world2_gloo_and_reshard(cpu_context, world):
    here = os.path.dirname(os.path.abspath(__file__))
    out = tempfile.mkdtemp() + "/out.pt"
    env = dict(os.environ, EXB_MP_OUT=out, OMP_NUM_THREADS="1")
    port = 29600 + os.getpid() % 200 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, "mp_cpu_check.py")],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "MP_CPU_CHECK_PASSED" in r.stdout, r.stdout[-3000:]
    saved = torch.load(out)
    import openembedding_b200.torch as embed
    from openembedding_b200.models.ctr import CriteoLR
    model = CriteoLR(num_sparse=26, num_dense=13, input_dim=1000000, num_shards=16)
    embed.load_server_model(model, saved["dir"] + "/ck")        # world=1 reads the world=2 checkpoint
    assert torch.equal(model.embeddings(torch.arange(0, 50)).detach(), saved["rows"])


def test_model_api_world2_gloo():
    """distributed_model save / load_weights / save_as_original_model on a 2-rank gloo job"""
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29800 + os.getpid() % 150
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, "mp_cpu_api_check.py")],
                       env=dict(os.environ, OMP_NUM_THREADS="1"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300)
    assert r.returncode == 0 and "MP_CPU_API_CHECK_PASSED" in r.stdout, r.stdout[-3000:]


def test_fewer_shards_than_ranks_world2_gloo():
    """num_shards=1 tables (array + hash) on a 2-rank job: round-robin placement, pulls agree, save/load"""
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29950 + os.getpid() % 40
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, "mp_cpu_shards_check.py")],
                       env=dict(os.environ, OMP_NUM_THREADS="1"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300)
    assert r.returncode == 0 and "SHARD1_OK" in r.stdout, r.stdout[-3000:]
