"""The examples double as end-to-end tests (reference: build.sh unit_test runs its examples)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    e.update(env or {})
    r = subprocess.run([sys.executable] + args, cwd=os.path.join(ROOT, "examples"), env=e, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def test_lr_checkpoint_reload():
    d = tempfile.mkdtemp()
    out = _run(["criteo_lr_subclass.py", "--epochs", "2", "--checkpoint", d + "/ck", "--save", d + "/saved"])
    assert "epoch 2" in out and os.path.exists(d + "/ck2.openembedding/openembedding/model_meta")
    out = _run(["criteo_lr_subclass.py", "--epochs", "1", "--load", d + "/ck2"])
    assert "epoch 1" in out


def test_deepctr_models_one_batch_edge_cases():
    for model, bs in (("DeepFM", 100), ("WDL", 50), ("xDeepFM", 10), ("DCN", 64), ("LR", 16)):
        out = _run(["criteo_deepctr_network.py", "--model", model, "--cpu", "--epochs", "1", "--batch_size", str(bs), "--cache"])
        assert "epoch 1" in out, model


def test_hook_then_serving():
    d = tempfile.mkdtemp()
    out = _run(["criteo_deepctr_hook.py"], env={"EXB_EXAMPLE_OUT": d + "/m"})
    assert "saved to" in out and os.path.exists(d + "/m/standalone.pt")
    out = _run(["serving_restful.py", "--model", d + "/m/openembedding", "--nodes", "2", "--port", "0"])
    assert "NORMAL" in out and "tensor" in out


def test_preprocess_tool():
    d = tempfile.mkdtemp()
    with open(d + "/day", "w") as fh:
        for r in range(20):
            fh.write("\t".join([str(r % 2)] + [str(r + i) for i in range(13)] + ["%x" % ((r * 7 + i) % 5) for i in range(26)]) + "\n")
    _run(["criteo_preprocess.py", "--in", d + "/day", "--out", d + "/out.csv", "--meta", d + "/meta", "--repeat", "2"])
    lines = open(d + "/out.csv").read().strip().split("\n")
    assert len(lines) == 41 and lines[0].startswith("label,I1") and len(lines[1].split(",")) == 40
    meta = dict(l.split() for l in open(d + "/meta"))
    assert int(meta["C1"]) == 10          # 5 distinct values x 2 repeats


def test_deepctr_world2_checkpoint_then_load_in_one_process():
    """horovodrun -np 2 ... --checkpoint, then load with a different worker count (reference build.sh:136-144)"""
    d = tempfile.mkdtemp()
    e = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    port = 29990 - os.getpid() % 30
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), "criteo_deepctr_network.py", "--model", "WDL", "--cpu",
                        "--epochs", "2", "--batch_size", "16", "--cache", "--checkpoint", d + "/ck"],
                       cwd=os.path.join(ROOT, "examples"), env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "epoch 2" in r.stdout, r.stdout[-3000:]
    assert os.path.exists(d + "/ck2/model_meta")
    out = _run(["criteo_deepctr_network.py", "--model", "WDL", "--cpu", "--epochs", "1", "--batch_size", "16", "--cache",
                "--load", d + "/ck2"])
    assert "epoch 1" in out
