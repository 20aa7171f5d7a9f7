"""Serving stack: export -> controller create_model (replicas) -> read-only pulls -> node
failure -> failover + restore. Reference analogues: c_api_test `model_mix`/`model_shard_num`
(openembedding/entry/c_api_test.cpp:40-66) and c_api_ha_test (kill servers while pulling)."""
import json
import tempfile
import time
import urllib.error
import urllib.request

import pytest
import torch


def _export_model(d):
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context
    ctx = get_context()
    emb = embed.Embedding(5000, 6, embeddings_initializer="uniform")
    hemb = embed.Embedding(-1, 3, embeddings_initializer="normal")
    opt = embed.distributed_optimizer(torch.optim.Adagrad(list(emb.parameters()) + list(hemb.parameters()), lr=0.1,
                                                          initial_accumulator_value=0.1))
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        x = torch.randint(0, 500, (64,), generator=g)
        loss = (emb(x).sum(-1) + hemb(x * 7919).sum(-1)).pow(2).mean()
        opt.zero_grad(); loss.backward(); opt.step()
    embed.save_server_model(None, d + "/model", include_optimizer=False)
    ids = torch.arange(0, 600)
    return ctx.model_sign(), emb(ids).detach().clone(), hemb(ids * 7919).detach().clone()


def test_master_tree_barrier_lock():
    from openembedding_b200.master import Master
    m = Master()
    a, b = m.client(), m.client()
    assert a.tree_node_add("x/y", "1") and not b.tree_node_add("x/y", "2")
    assert b.tree_node_get("x/y") == "1"
    a.tree_node_set("x/z", "3")
    assert sorted(b.tree_node_sub("x")) == ["y", "z"]
    assert a.tree_node_del("x/y") and b.tree_node_get("x/y") is None and b.tree_node_sub("x") == ["z"]
    assert a.tree_node_add("x/y", "again")
    assert [a.generate_id("n"), b.generate_id("n"), a.generate_id("n")] == [0, 1, 2]
    a.acquire_lock("L")
    with pytest.raises(TimeoutError):
        b.acquire_lock("L", timeout=0.1)
    a.release_lock("L")
    b.acquire_lock("L", timeout=1)
    import threading
    done = []
    ts = [threading.Thread(target=lambda c=c: (c.barrier("B", 2), done.append(1))) for c in (a, b)]
    [t.start() for t in ts]
    [t.join(5) for t in ts]
    assert done == [1, 1]


@pytest.mark.parametrize("shard_num,replicas", [(-1, 2), (7, 3), (1, 1)])
def test_serving_replicas_failover_restore(cpu_context, shard_num, replicas):
    import openembedding_b200 as oe
    from openembedding_b200.serving.client import NoReplica, ServingClient
    from openembedding_b200.serving.controller import ModelController, serve
    d = tempfile.mkdtemp()
    sign, want_e, want_h = _export_model(d)
    master = oe.Master()
    servers = [oe.Server(master_endpoint=master.endpoint) for _ in range(3)]
    httpd, _ = serve(master.endpoint, port=0, bind_ip="127.0.0.1")
    import threading
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    base = "http://127.0.0.1:%d" % httpd.server_address[1]
    req = urllib.request.Request(base + "/models", method="POST", headers={"Content-Type": "application/json"},
                                 data=json.dumps({"model_uri": d + "/model", "replica_num": replicas,
                                                  "shard_num": shard_num}).encode())
    got = json.loads(urllib.request.urlopen(req, timeout=60).read())
    assert got["model_sign"] == sign
    rec = json.loads(urllib.request.urlopen(base + "/models/" + sign).read())
    assert rec["model_status"] == "NORMAL" and rec["replica_num"] == min(replicas, 3)
    assert len(json.loads(urllib.request.urlopen(base + "/nodes").read())) == 3
    cli = ServingClient(master.endpoint)
    ids = torch.arange(0, 600)
    ve, vh = cli.find_model_variable(sign, 0), cli.find_model_variable(sign, 1)
    assert torch.allclose(ve.pull(ids), want_e) and torch.allclose(vh.pull(ids * 7919), want_h)
    if replicas >= 2:
        # kill one node: pulls keep working from the other replica
        servers[0].exit()
        time.sleep(0.2)
        assert torch.allclose(ve.pull(ids, timeout=20), want_e)
        # a fresh node restores the dead node's shards (peer streaming) and takes its place
        fresh = oe.Server(master_endpoint=master.endpoint)
        restored = ModelController(master.endpoint).restore_node(fresh.node_id, fresh.endpoint)
        assert restored
        deadline = time.time() + 20
        while time.time() < deadline:
            st = json.loads(urllib.request.urlopen("http://%s/models" % fresh.endpoint).read())
            if st.get(sign, {}).get("status") == "NORMAL":
                break
            time.sleep(0.05)
        servers[1].exit()        # now the ORIGINAL second replica dies too
        time.sleep(0.2)
        cli2 = ServingClient(master.endpoint)
        assert torch.allclose(cli2.find_model_variable(sign, 0).pull(ids, timeout=20), want_e)
        servers.append(fresh)
    else:
        servers[0].exit(); servers[1].exit(); servers[2].exit()
        with pytest.raises(NoReplica):
            ve.pull(ids, timeout=1.0)
    urllib.request.urlopen(urllib.request.Request(base + "/models/" + sign, method="DELETE")).read()
    for s in servers:
        try:
            s.exit()
        except Exception:
            pass
    httpd.shutdown()


def test_daemons_end_to_end(cpu_context):
    """the deployment of the reference's run/*.sh: masterd + 2 server daemons + controller daemon as separate
    processes, model created over REST, pulled through the client"""
    import os
    import socket
    import subprocess
    import sys
    import openembedding_b200.torch as embed
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context
    from openembedding_b200.master import MasterClient
    from openembedding_b200.serving.client import ServingClient
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    v = embed.Variable(shape=(500, 4), name="v", num_shards=1, initializer={"category": "uniform", "minval": -1.0, "maxval": 1.0})
    ids = torch.arange(0, 500, 5)
    v.push_gradients(ids, torch.ones(ids.numel(), 4))
    v.update_weights()
    want = v.sparse_read(torch.arange(40)).clone()
    d = tempfile.mkdtemp()
    checkpoint.save_model(get_context(), d + "/m", include_optimizer=False)
    sign = get_context().model_sign()
    procs = []
    try:
        m = subprocess.Popen([sys.executable, "-m", "openembedding_b200.master", "--port", "0"], cwd=root,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(m)
        endpoint = m.stdout.readline().split()[-1]
        for _ in range(2):
            procs.append(subprocess.Popen([sys.executable, "-m", "openembedding_b200.serving.node", "--master_endpoint", endpoint],
                                          cwd=root, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        mc = MasterClient(endpoint)
        t0 = time.time()
        while len(mc.tree_node_sub("nodes")) < 2:
            assert time.time() - t0 < 60, "server daemons did not register"
            time.sleep(0.1)
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        procs.append(subprocess.Popen([sys.executable, "-m", "openembedding_b200.serving.controller", "--master_endpoint", endpoint,
                                       "--port", str(port), "--bind_ip", "127.0.0.1"], cwd=root,
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        base = "http://127.0.0.1:%d" % port
        t0 = time.time()
        while True:
            try:
                urllib.request.urlopen(base + "/nodes", timeout=2).read()
                break
            except Exception:
                assert time.time() - t0 < 60, "controller daemon did not come up"
                time.sleep(0.2)
        req = urllib.request.Request(base + "/models", method="POST", headers={"Content-Type": "application/json"},
                                     data=json.dumps({"model_uri": d + "/m", "replica_num": 2, "shard_num": -1}).encode())
        assert json.loads(urllib.request.urlopen(req, timeout=120).read())["model_sign"] == sign
        assert json.loads(urllib.request.urlopen(base + "/models/" + sign).read())["model_status"] == "NORMAL"
        got = ServingClient(endpoint).find_model_variable(sign, 0).pull(torch.arange(40))
        assert torch.allclose(got, want)
        # REST introspection and teardown (controller.cc: GET/DELETE /models[/sign], GET/DELETE /nodes[/id])
        assert sign in json.loads(urllib.request.urlopen(base + "/models").read())
        nodes = json.loads(urllib.request.urlopen(base + "/nodes").read())
        assert len(nodes) == 2
        nid = sorted(nodes)[0]
        assert json.loads(urllib.request.urlopen(base + "/nodes/" + nid).read())["node_id"] == int(nid)
        req = urllib.request.Request(base + "/models/" + sign, method="DELETE")
        assert json.loads(urllib.request.urlopen(req).read())["deleted"]
        with pytest.raises(urllib.error.HTTPError):
            urllib.request.urlopen(base + "/models/" + sign)
        req = urllib.request.Request(base + "/nodes/" + nid, method="DELETE")
        assert json.loads(urllib.request.urlopen(req).read())["shutdown"]
        t0 = time.time()
        while len(json.loads(urllib.request.urlopen(base + "/nodes").read())) != 1:
            assert time.time() - t0 < 30
            time.sleep(0.2)
    finally:
        for p in reversed(procs):
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()


def test_daemon_sigkill_failover_and_restore(cpu_context):
    """reference c_api_ha_test: SIGKILL a server process while a client pulls; pulls keep succeeding from the
    replica; a replacement daemon started with --restore takes over ALL shard replicas of the dead node"""
    import os
    import signal
    import subprocess
    import sys
    import openembedding_b200.torch as embed
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context
    from openembedding_b200.master import MasterClient
    from openembedding_b200.serving.client import ServingClient
    from openembedding_b200.serving.controller import ModelController
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    v = embed.Variable(shape=(600, 4), name="v", num_shards=1, initializer={"category": "uniform", "minval": -1.0, "maxval": 1.0})
    ids = torch.arange(0, 600, 2)
    v.push_gradients(ids, torch.ones(ids.numel(), 4))
    v.update_weights()
    probe = torch.arange(60)
    want = v.sparse_read(probe).clone()
    d = tempfile.mkdtemp()
    checkpoint.save_model(get_context(), d + "/m", include_optimizer=False)
    procs = []

    def node(*extra):
        p = subprocess.Popen([sys.executable, "-m", "openembedding_b200.serving.node", "--master_endpoint", endpoint] + list(extra),
                             cwd=root, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        procs.append(p)
        return p

    def wait_nodes(n):
        t0 = time.time()
        while len(mc.tree_node_sub("nodes")) != n:
            assert time.time() - t0 < 60, "expected %d live nodes, have %s" % (n, mc.tree_node_sub("nodes"))
            time.sleep(0.1)
    try:
        m = subprocess.Popen([sys.executable, "-m", "openembedding_b200.master", "--port", "0"], cwd=root,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(m)
        endpoint = m.stdout.readline().split()[-1]
        mc = MasterClient(endpoint)
        victims = [node() for _ in range(3)]
        wait_nodes(3)
        ctl = ModelController(endpoint)
        sign = ctl.create_model(d + "/m", replica_num=2, shard_num=3)
        cli = ServingClient(endpoint)
        var = cli.find_model_variable(sign, 0)
        assert torch.allclose(var.pull(probe), want)
        victims[0].send_signal(signal.SIGKILL)                 # no clean shutdown: the lease has to expire
        victims[0].wait(timeout=10)
        for _ in range(5):                                     # pulls fail over to the surviving replicas
            assert torch.allclose(var.pull(probe, timeout=30), want)
        wait_nodes(2)
        node("--restore")
        wait_nodes(3)
        t0 = time.time()
        while True:
            rec = ctl.show_model(sign)
            live = set(ctl.nodes())
            if all(set(reps) <= live for reps in rec["placement"].values()):
                break
            assert time.time() - t0 < 60, rec["placement"]
            time.sleep(0.2)
        assert all(len(set(reps)) == 2 for reps in rec["placement"].values())
        assert torch.allclose(ServingClient(endpoint).find_model_variable(sign, 0).pull(probe), want)
    finally:
        for p in reversed(procs):
            if p.poll() is None:
                p.terminate()
        for p in procs:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()


def test_placement_cursor_and_hash_policy(cpu_context):
    """single-shard models rotate over the nodes (Model.cpp:153-186 rotating cursor, cluster-wide in the master);
    ServingClient(policy="hash") keeps a key on one replica (jump consistent hash)"""
    import openembedding_b200 as oe
    from openembedding_b200.serving.client import ServingClient
    from openembedding_b200.serving.controller import ModelController
    d = tempfile.mkdtemp()
    sign, want_e, _ = _export_model(d)
    master = oe.Master()
    servers = [oe.Server(master_endpoint=master.endpoint) for _ in range(3)]
    ctl = ModelController(master.endpoint)
    import shutil
    firsts, signs = [], []
    for k in range(3):                       # the same files under three signs = three single-shard models
        dk = d + "/copy%d" % k
        shutil.copytree(d + "/model", dk)
        meta = json.load(open(dk + "/model_meta"))
        meta["model_sign"] = "%s-copy%d" % (sign, k)
        json.dump(meta, open(dk + "/model_meta", "w"), indent=4)
        s = ctl.create_model(dk, replica_num=2, shard_num=1)
        signs.append(s)
        reps = ctl.show_model(s)["placement"]["0"]
        assert len(reps) == 2 and len(set(reps)) == 2
        firsts.append(reps[0])
    assert len(set(firsts)) == 3, firsts      # cursor advanced by shard_num * replica_num = 2 on a ring of 3
    # replicas on two nodes: the hash policy sends one key to one replica, different keys spread over both
    s = signs[0]
    cli = ServingClient(master.endpoint, policy="hash")
    reps = cli._model(s)["placement"]["0"]
    picks = [cli._pick(reps, key=k) for k in range(200)]
    assert all(cli._pick(reps, key=k) == picks[k] for k in range(200))
    assert len(set(picks)) == len(reps)
    v = cli.find_model_variable(s, 0)
    assert torch.allclose(v.pull(torch.arange(0, 600)), want_e)
    for sv in servers:
        sv.exit()
