"""Config tiers and the byte-exact checkpoint format."""
import json
import os
import struct
import tempfile

import numpy as np
import pytest
import torch
import yaml


def test_env_config_schema():
    from openembedding_b200.config import EnvConfig
    c = EnvConfig('{"server":{"server_concurrency":28, "cache_size":500}}')
    assert c["server"]["server_concurrency"] == 28 and c["server"]["cache_size"] == 500
    assert c["server"]["update_early_return"] is True and c["master"]["root_path"] == "/openembedding"
    EnvConfig("server:\n  message_compress: lz4\n")
    with pytest.raises(ValueError):
        EnvConfig('{"server":{"no_such_key":1}}')
    with pytest.raises(ValueError):
        EnvConfig('{"server":{"message_compress":"brotli"}}')
    assert "rpc" in yaml.safe_load(c.dump_yaml())


def test_optimizer_initializer_configs():
    from openembedding_b200.config import (dump_variable_config, load_variable_config, normalize_initializer,
                                           normalize_optimizer, optimizer_state_dim)
    o = normalize_optimizer({"category": "Adam", "learning_rate": "0.01", "unknown": 3})
    assert o == {"category": "adam", "learning_rate": 0.01, "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-7}
    assert optimizer_state_dim(o, 8) == 2 * 8 + 2 and optimizer_state_dim({"category": "adamax"}, 8) == 17
    assert optimizer_state_dim({"category": "adagrad"}, 9) == 9 and optimizer_state_dim({"category": "default"}, 9) == 0
    with pytest.raises(ValueError):
        normalize_optimizer({"category": "nadam"})
    assert normalize_initializer("uniform") == {"category": "uniform", "minval": -0.05, "maxval": 0.05}
    assert normalize_initializer("zeros") == {"category": "constant", "value": 0.0}
    with pytest.raises(ValueError):
        normalize_initializer("glorot_uniform")
    assert normalize_initializer("glorot_uniform", explicit=False)["category"] == "constant"
    text = dump_variable_config("hash", 123, {"category": "ftrl", "learning_rate": 0.3}, "normal")
    doc = yaml.safe_load(text)
    assert doc["table"] == "hash" and doc["reserve_items"] == 123 and doc["optimizer"] == "ftrl"
    assert doc["ftrl"]["learning_rate"] == 0.3 and doc["initializer"] == "normal" and doc["normal"]["stddev"] == 0.05
    back = load_variable_config(text)
    assert back["optimizer"]["category"] == "ftrl" and back["optimizer"]["learning_rate_power"] == -0.5
    assert "optimizer" not in yaml.safe_load(dump_variable_config("array", 1, {"category": "sgd"}, "zeros",
                                                                  include_optimizer=False))


def _parse_shard_file(path):
    """independent parser written from the format description (NOT the package's reader)"""
    recs = []
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        vid, dtype, dim, vocab, clen = struct.unpack_from("<IiQQQ", data, off)
        off += 4 + 4 + 8 + 8 + 8
        cfg = data[off:off + clen].decode()
        off += clen
        shard_id, shard_num, sls, nitems = struct.unpack_from("<iiQQ", data, off)
        off += 4 + 4 + 8 + 8
        rows = []
        done = 0
        while done < nitems:
            (n,) = struct.unpack_from("<Q", data, off)
            off += 8
            idx = np.frombuffer(data, dtype="<u8", count=n, offset=off)
            off += 8 * n
            itemsize = 4 if dtype == 0x104 else 8
            w = np.frombuffer(data, dtype="<f4" if itemsize == 4 else "<f8", count=n * dim, offset=off).reshape(n, dim)
            off += n * dim * itemsize
            st = np.frombuffer(data, dtype=np.uint8, count=n * sls, offset=off).reshape(n, sls)
            off += n * sls
            rows.append((idx, w, st))
            done += n
        recs.append(dict(vid=vid, dtype=dtype, dim=dim, vocab=vocab, cfg=cfg, shard_id=shard_id, shard_num=shard_num,
                         sls=sls, nitems=nitems, rows=rows))
    return recs


@pytest.mark.parametrize("include_optimizer", [True, False])
def test_checkpoint_bytes(cpu_context, include_optimizer):
    import openembedding_b200.torch as embed
    from openembedding_b200.context import get_context
    ctx = get_context()
    e = embed.Embedding(300, 5, embeddings_initializer="uniform")
    h = embed.Embedding(-1, 2, embeddings_initializer="zeros", dtype=torch.float64)
    opt = embed.distributed_optimizer(torch.optim.Adam(list(e.parameters()) + list(h.parameters()), lr=0.01))
    x = torch.tensor([5, 9, 9, 250])
    loss = (e(x).sum() + h(x * 10 ** 12).sum().float()) ** 2
    loss.backward(); opt.step()
    d = tempfile.mkdtemp() + "/m"
    embed.save_server_model(None, d, include_optimizer=include_optimizer)
    meta = json.load(open(d + "/model_meta"))
    assert list(meta.keys()) == ["model_sign", "variables", "version"] and meta["version"] == "0.2"
    assert meta["model_sign"] == ctx.model_sign() and meta["model_sign"].endswith("-1")
    assert meta["variables"][0] == {"datatype": "float32", "embedding_dim": 5, "vocabulary_size": 300, "storage_name": "0"}
    assert meta["variables"][1]["vocabulary_size"] == 2 ** 63 and meta["variables"][1]["datatype"] == "float64"
    assert open(d + "/model_meta").read() == json.dumps(meta, indent=4)
    r0 = _parse_shard_file(d + "/0/model_0_0")[0]
    assert (r0["vid"], r0["dtype"], r0["dim"], r0["vocab"], r0["shard_id"], r0["shard_num"]) == (0, 0x104, 5, 300, 0, 1)
    assert r0["nitems"] == 3 and r0["sls"] == ((2 * 5 + 2) * 4 if include_optimizer else 0)
    idx = np.concatenate([b[0] for b in r0["rows"]])
    assert sorted(idx.tolist()) == [5, 9, 250]
    cfg = yaml.safe_load(r0["cfg"])
    assert cfg["table"] == "array" and cfg["initializer"] == "uniform" and (("adam" in cfg) == include_optimizer)
    w = np.concatenate([b[1] for b in r0["rows"]])
    want = e(torch.from_numpy(idx.astype(np.int64))).detach().numpy()
    assert np.array_equal(w, want)
    if include_optimizer:
        st = np.concatenate([b[2] for b in r0["rows"]]).view("<f4").reshape(3, 12)
        assert np.allclose(st[:, 10], 0.9) and np.allclose(st[:, 11], 0.999)      # per-row beta powers after 1 step
    r1 = _parse_shard_file(d + "/1/model_0_0")[0]
    assert r1["dtype"] == 0x108 and r1["vocab"] == 2 ** 63 and yaml.safe_load(r1["cfg"])["table"] == "hash"
    assert sorted(np.concatenate([b[0] for b in r1["rows"]]).tolist()) == [5 * 10 ** 12, 9 * 10 ** 12, 250 * 10 ** 12]


def test_load_reshards_foreign_shard_count(cpu_context):
    """a checkpoint written by 3 shards loads into a 1-shard job (re-hash on load)"""
    import openembedding_b200.torch as embed
    from openembedding_b200 import _native
    from openembedding_b200.config import dump_variable_config
    e = embed.Embedding(100, 2, embeddings_initializer="zeros")
    from openembedding_b200.context import get_context
    ctx = get_context()
    d = tempfile.mkdtemp() + "/m"
    os.makedirs(d + "/0")
    meta = {"model_sign": "x-0", "variables": [{"datatype": "float32", "embedding_dim": 2, "vocabulary_size": 100,
                                                "storage_name": "0"}], "version": "0.2"}
    open(d + "/model_meta", "w").write(json.dumps(meta, indent=4))
    lib = _native.core()
    cfg = dump_variable_config("array", 0, {"category": "default"}, "zeros").encode()
    for shard in range(3):
        w = lib.exb_fw_open((d + "/0/model_%d_0" % shard).encode())
        ids = np.array([g for g in range(100) if g % 3 == shard and g % 7 == 0], dtype=np.uint64)
        local = np.ascontiguousarray(ids // 3)
        rows = np.ascontiguousarray(np.stack([ids, ids * 2], 1).astype(np.float32))
        lib.exb_fw_header(w, 0, 0x104, 2, 100, cfg, len(cfg), shard, 3, 0, ids.size)
        lib.exb_fw_block(w, ids.size, local.ctypes.data, rows.ctypes.data, rows.nbytes, None, 0)
        lib.exb_fw_close(w)
    embed.load_server_model(None, d)
    got = e(torch.arange(100)).detach()
    want = torch.zeros(100, 2)
    for g in range(0, 100, 7):
        want[g] = torch.tensor([g, 2 * g], dtype=torch.float32)
    assert torch.equal(got, want)
    bad = dict(meta, version="0.1")
    open(d + "/model_meta", "w").write(json.dumps(bad))
    with pytest.raises(ValueError):
        embed.load_server_model(None, d)


def test_checkpoint_multiple_dump_files_and_optimizer_change(cpu_context):
    """server.server_dump_files > 1 spreads the shards over files (shard_id % files); loading with a different
    optimizer category keeps the weights and resets the state (EmbeddingVariable.cpp:44-47)"""
    import os
    import tempfile
    import openembedding_b200 as oe
    import openembedding_b200.torch as embed
    from openembedding_b200 import checkpoint
    from openembedding_b200.context import get_context, reset_context
    reset_context()
    old = oe.flags.config
    oe.flags.config = '{"server": {"server_dump_files": 3}}'
    try:
        ctx = get_context()
        emb = embed.Embedding(1000, 4, embeddings_initializer={"category": "uniform", "minval": -1.0, "maxval": 1.0})
        opt = embed.distributed_optimizer(torch.optim.Adagrad(emb.parameters(), lr=0.1, initial_accumulator_value=0.1))
        ids = torch.arange(0, 1000, 3)
        for _ in range(3):
            loss = (emb(ids) ** 2).sum()
            opt.zero_grad(); loss.backward(); opt.step()
        want = emb(torch.arange(1000)).detach().clone()
        d = tempfile.mkdtemp()
        checkpoint.save_model(ctx, d + "/ck", include_optimizer=True)
        files = sorted(os.listdir(d + "/ck/0"))
        assert files and all(f.startswith("model_0_") for f in files) and len(files) <= 3
        # reload into a fresh job that uses a different optimizer: weights survive, state starts over
        reset_context()
        ctx = get_context()
        emb2 = embed.Embedding(1000, 4, embeddings_initializer={"category": "constant", "value": 0.0})
        opt2 = embed.distributed_optimizer(torch.optim.Adam(emb2.parameters(), lr=0.01))
        checkpoint.load_model(ctx, d + "/ck")
        assert torch.equal(emb2(torch.arange(1000)).detach(), want)
        loss = (emb2(ids) ** 2).sum()
        opt2.zero_grad(); loss.backward(); opt2.step()          # Adam step from zero moments: |dw| == lr
        moved = (emb2(ids).detach() - want[ids]).abs()
        assert torch.allclose(moved[want[ids].abs() > 1e-3], torch.tensor(0.01), atol=2e-4)
    finally:
        oe.flags.config = old
        reset_context()


def test_load_skips_foreign_segments_and_config_only(cpu_context):
    """a loader reads only the segments whose shard it owns (the others are seeked past); restore_config_only restores
    initializer / optimizer configs without loading a row"""
    import openembedding_b200.torch as embed
    from openembedding_b200 import _native, checkpoint
    from openembedding_b200.config import dump_variable_config
    from openembedding_b200.context import get_context, reset_context
    lib = _native.core()
    d = tempfile.mkdtemp()
    fn = d + "/model_0_0"
    cfg = dump_variable_config("array", 0, {"category": "default"}, "zeros").encode()
    w = lib.exb_fw_open(fn.encode())
    for shard in range(3):                       # three segments (saved shards) in one file
        local = np.arange(5, dtype=np.uint64)
        rows = np.full((5, 2), float(shard), dtype=np.float32)
        lib.exb_fw_header(w, 0, 0x104, 2, 100, cfg, len(cfg), shard, 3, 0, local.size)
        lib.exb_fw_block(w, local.size, local.ctypes.data, rows.ctypes.data, rows.nbytes, None, 0)
    lib.exb_fw_close(w)
    recs = list(checkpoint.iter_shard_file(fn, want=lambda h: h["shard_id"] == 1))
    assert [r[0] for r in recs] == ["header", "header", "block", "header"]
    blk = [r for r in recs if r[0] == "block"][0]
    assert blk[1]["shard_id"] == 1 and (blk[3] == 1.0).all() and list(blk[2]) == [1, 4, 7, 10, 13]
    assert [r[0] for r in checkpoint.iter_shard_file(fn, want=lambda h: False)] == ["header"] * 3
    assert sum(r[0] == "block" for r in checkpoint.iter_shard_file(fn)) == 3

    reset_context()
    ctx = get_context()
    emb = embed.Embedding(50, 2, embeddings_initializer={"category": "constant", "value": 0.5})
    opt = embed.distributed_optimizer(torch.optim.Adagrad(emb.parameters(), lr=0.1, initial_accumulator_value=0.1))
    loss = (emb(torch.arange(10)) ** 2).sum()
    opt.zero_grad(); loss.backward(); opt.step()
    checkpoint.save_model(ctx, d + "/ck", include_optimizer=True)
    reset_context()
    ctx = get_context()
    emb2 = embed.Embedding(50, 2, embeddings_initializer={"category": "constant", "value": 0.0})
    checkpoint.load_model(ctx, d + "/ck", restore_config_only=True)
    var = emb2.variable.variable
    assert var.optimizer["category"] == "adagrad" and abs(var.optimizer["learning_rate"] - 0.1) < 1e-9
    assert torch.equal(emb2(torch.arange(50)).detach(), torch.full((50, 2), 0.5))   # restored initializer, no rows
    reset_context()
