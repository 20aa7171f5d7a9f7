"""Status vocabulary, URI/file staging, DataType registry, message compression, masterd entry."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import pytest
import torch

from openembedding_b200.config import DataType
from openembedding_b200.status import Status, StatusError, check, retry
from openembedding_b200.utils.fs import Staging, URIConfig, exists


def test_status_retry():
    assert Status.OK.ok and Status.NO_REPLICA.retryable and not Status.FATAL.retryable
    with pytest.raises(StatusError):
        check(Status.INVALID_ID, "bad id")
    calls = []

    def flaky():
        calls.append(1)
        if len(calls) < 3:
            raise StatusError(Status.TIMEOUT)
        return 7
    refreshed = []
    assert retry(flaky, attempts=3, refresh=lambda: refreshed.append(1)) == 7 and len(refreshed) == 2
    with pytest.raises(StatusError):
        retry(lambda: check(Status.FATAL), attempts=3)


def test_datatype():
    dt = DataType("float32")
    assert dt.size == 4 and int(dt) == 0x104 and dt.is_table_type and dt == "float32" and dt == 0x104
    assert DataType(torch.float64).size == 8 and DataType(0x8).name == "int64" and not DataType("int8").is_table_type
    with pytest.raises(ValueError):
        DataType("float16")


def test_uri_and_staging():
    u = URIConfig("hdfs://nn/models/a?format=archive&x=1")
    assert u.scheme == "hdfs" and u.path == "hdfs://nn/models/a" and u.params == {"format": "archive", "x": "1"}
    assert URIConfig("/tmp/x").is_local and URIConfig("file:///tmp/x").path == "/tmp/x"
    assert URIConfig("mem://null/").is_null
    d = tempfile.mkdtemp()
    with Staging(d + "/m", "w") as p:
        open(os.path.join(p, "f"), "w").write("1")
    assert exists(d + "/m/f")
    with Staging("mem://null/", "w") as p:
        open(os.path.join(p, "f"), "w").write("1")
    assert not os.path.exists(p)


def test_checkpoint_through_file_uri(cpu_context):
    import openembedding_b200.torch as embed
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context
    v = embed.Variable(shape=(100, 4), name="v", num_shards=1)
    ctx = get_context()
    ids = torch.arange(10)
    v.push_gradients(ids, torch.ones(10, 4))
    v.update_weights()
    d = tempfile.mkdtemp()
    checkpoint.save_model(ctx, "file://" + d + "/ck")
    before = v.sparse_read(ids).clone()
    v.push_gradients(ids, torch.ones(10, 4))
    v.update_weights()
    checkpoint.load_model(ctx, "file://" + d + "/ck")
    assert torch.equal(v.sparse_read(ids), before)


def test_serving_message_compress():
    import openembedding_b200 as oe
    import openembedding_b200.torch as embed
    from openembedding_b200.context import reset_context
    from openembedding_b200.serving.client import ServingClient
    from openembedding_b200.serving.controller import ModelController
    reset_context()
    oe.flags.device = "cpu"
    v = embed.Variable(shape=(5000, 8), name="v", num_shards=1,
                       initializer={"category": "uniform", "minval": -1.0, "maxval": 1.0})
    ids = torch.arange(3000)
    ref = v.sparse_read(ids).clone()
    v.push_gradients(ids, torch.zeros(3000, 8))
    v.update_weights()
    d = tempfile.mkdtemp()
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context
    checkpoint.save_model(get_context(), d + "/m", include_optimizer=False)
    sign = get_context().model_sign()
    master = oe.Master()
    node = oe.Server(master_endpoint=master.endpoint)
    ctl = ModelController(master.endpoint)
    sign = ctl.create_model(d + "/m", replica_num=1, shard_num=1)
    out_plain = ServingClient(master.endpoint).find_model_variable(sign, 0).pull(ids)
    out_z = ServingClient(master.endpoint, message_compress="zlib").find_model_variable(sign, 0).pull(ids)
    assert torch.allclose(out_plain, ref) and torch.equal(out_plain, out_z)
    for codec in ("lz4", "snappy"):        # native LZ4 block codec (utils/compress.py) in both directions
        out_l = ServingClient(master.endpoint, message_compress=codec).find_model_variable(sign, 0).pull(ids)
        assert torch.equal(out_plain, out_l), codec
    node.exit()
    reset_context()


def test_masterd_entry():
    p = subprocess.Popen([sys.executable, "-m", "openembedding_b200.master", "--port", "0"], stdout=subprocess.PIPE,
                         text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        line = p.stdout.readline()
        assert line.startswith("master endpoint ")
        from openembedding_b200.master import MasterClient
        c = MasterClient(line.split()[-1])
        assert c.tree_node_add("x", "1") and c.tree_node_get("x") == "1" and c.generate_id("n") == 0
    finally:
        p.terminate()
        p.wait(timeout=10)


def test_report_interval_enables_timers_and_counters(capsys):
    """server.report_interval > 0 -> vtimers / pull_indices / pull_unique are collected, Monitor prints them"""
    import openembedding_b200 as oe
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context, reset_context
    from openembedding_b200.utils import timers
    reset_context()
    timers.reset()
    oe.flags.device = "cpu"
    old = oe.flags.config
    oe.flags.config = '{"server": {"report_interval": 1}}'
    try:
        v = embed.Variable(shape=(100, 4), name="v", num_shards=1)
        ids = torch.tensor([1, 2, 2, 3, 3, 3])
        v.sparse_read(ids)
        v.push_gradients(ids, torch.ones(6, 4))
        v.update_weights()
        snap = timers.snapshot(get_context())
        assert snap["counters"]["pull_indices"] >= 6 and snap["counters"]["pull_unique"] >= 3
        assert "client.pull_weights" in snap["timers"] and "client.update_weights" in snap["timers"]
        time.sleep(1.4)
        assert "client.pull_weights" in capsys.readouterr().out          # the Monitor thread printed the table
    finally:
        oe.flags.config = old
        reset_context()
        timers.enable(False)
        timers.reset()


def test_checkpoint_through_hdfs_pipe(cpu_context, tmp_path, monkeypatch):
    """hdfs:// models are staged through `hdfs dfs -put/-get` (reference: ShellUtility pipes). A fake `hdfs`
    executable that maps hdfs://nn/<p> onto a local directory stands in for the cluster."""
    import stat
    import openembedding_b200.torch as embed
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context
    root = tmp_path / "fake_hdfs"
    root.mkdir()
    fake = tmp_path / "hdfs"
    fake.write_text("""#!/bin/bash
# usage: hdfs dfs -mkdir -p P | -put -f SRC... P | -get P/* DST | -test -e P
ROOT="%s"
map() { echo "$ROOT/${1#hdfs://nn/}"; }
shift   # dfs
case "$1" in
  -mkdir) mkdir -p "$(map "$3")" ;;
  -put) shift; shift; args=("$@"); dst="$(map "${args[-1]}")"; unset 'args[-1]'; mkdir -p "$dst"; cp -r "${args[@]}" "$dst"/ ;;
  -get) src="$(map "${2%%/\\*}")"; cp -r "$src"/* "$3"/ ;;
  -test) test -e "$(map "$3")" ;;
  *) exit 2 ;;
esac
""" % root)
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("HADOOP_BIN", str(fake))
    v = embed.Variable(shape=(100, 4), name="v", num_shards=1)
    ctx = get_context()
    ids = torch.arange(10)
    v.push_gradients(ids, torch.ones(10, 4))
    v.update_weights()
    before = v.sparse_read(ids).clone()
    checkpoint.save_model(ctx, "hdfs://nn/models/m1")
    assert (root / "models" / "m1" / "model_meta").exists()
    from openembedding_b200.utils.fs import exists
    assert exists("hdfs://nn/models/m1") and not exists("hdfs://nn/models/none")
    v.push_gradients(ids, torch.ones(10, 4))
    v.update_weights()
    checkpoint.load_model(ctx, "hdfs://nn/models/m1")
    assert torch.equal(v.sparse_read(ids), before)


def test_hashing_partitioner():
    from openembedding_b200.utils.hashing import Partitioner, jump_consistent_hash, murmur3_32, murmur3_fmix64
    assert murmur3_32(b"") == 0 and murmur3_32(b"", 1) == 0x514E28B7 and murmur3_32(b"hello") == 0x248BFA47
    assert murmur3_fmix64(0) == 0 and len({murmur3_fmix64(i) for i in range(1000)}) == 1000
    keys = list(range(20000))
    a = [jump_consistent_hash(murmur3_fmix64(k), 8) for k in keys]
    b = [jump_consistent_hash(murmur3_fmix64(k), 9) for k in keys]
    assert set(a) == set(range(8)) and max(a.count(i) for i in range(8)) < 20000 / 8 * 1.1
    moved = sum(1 for x, y in zip(a, b) if x != y)
    assert all(y == 8 for x, y in zip(a, b) if x != y) and 0.08 < moved / 20000 < 0.14    # ~1/9 of the keys move
    p = Partitioner(5)
    assert p("model-a") == p(b"model-a") and 0 <= p(12345) < 5


def test_native_model_in_process(cpu_context):
    """NativePS counterpart: load an exported model into this process and pull without any server"""
    import openembedding_b200.torch as embed
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context
    from openembedding_b200.serving.native import NativeModel
    v = embed.Variable(shape=(1000, 6), name="v", num_shards=1,
                       initializer={"category": "uniform", "minval": -1.0, "maxval": 1.0})
    ids = torch.arange(0, 300, 3)
    v.push_gradients(ids, torch.ones(ids.numel(), 6))
    v.update_weights()
    probe = torch.tensor([[0, 3, 6], [1, 2, 999]])        # trained and never-trained rows
    want = v.sparse_read(probe).clone()
    d = tempfile.mkdtemp()
    checkpoint.save_model(get_context(), d + "/m", include_optimizer=False)
    for shards in (1, 3):
        m = NativeModel(d + "/m", shard_num=shards)
        got = m.pull(0, probe)
        assert got.shape == (2, 3, 6) and torch.allclose(got, want)
        m.close()


def test_master_watchers():
    """MasterClient.watch: a change of a node or of one of its children wakes the watcher (server-side wait on the
    next version's notification key, no polling loop); reference: pico-core rpc/Master.cpp watchers"""
    import threading
    import time
    import openembedding_b200 as oe
    master = oe.Master()
    a, b = master.client(), master.client()
    seen, ev = [], threading.Event()

    def cb(path, ver):
        seen.append((path, ver))
        ev.set()

    w = a.watch("jobs", cb, poll_s=5.0)               # long wait: a wake-up within 2 s proves it was notified
    v0 = a.node_version("jobs")
    t0 = time.time()
    assert b.tree_node_add("jobs/j1", "x")
    assert ev.wait(2.0) and time.time() - t0 < 2.0, seen
    assert seen[-1][0] == "jobs" and seen[-1][1] > v0
    ev.clear()
    b.tree_node_set("jobs/j1", "y")                  # a child's value change notifies the parent's watchers too
    assert ev.wait(2.0)
    ev.clear()
    b.tree_node_del("jobs/j1")
    assert ev.wait(2.0)
    n = len(seen)
    w.cancel()
    b.tree_node_add("jobs/j2", "z")
    time.sleep(0.3)
    assert len(seen) == n                            # cancelled watchers stay quiet
    # a burst of changes is coalesced, never lost: the watcher ends at the current version
    seen2 = []
    w2 = a.watch("burst", lambda p, v: seen2.append(v), poll_s=0.2)
    for i in range(20):
        b.tree_node_set("burst/k%d" % i, str(i))
    deadline = time.time() + 3.0
    while time.time() < deadline and (not seen2 or seen2[-1] != a.node_version("burst")):
        time.sleep(0.05)
    assert seen2 and seen2[-1] == a.node_version("burst") == 20
    w2.cancel()
    a.close(); b.close()
