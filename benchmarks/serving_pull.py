"""Read-only pull throughput of the serving stack (master + nodes + client) -- counterpart of the
reference's test/benchmark/server.py (a standalone server for remote-PS runs) for the serving tier.

    python benchmarks/serving_pull.py --rows 200000 --dim 64 --nodes 2 --replicas 2 --batch 4096
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openembedding_b200 as oe  # noqa: E402
import openembedding_b200.torch as embed  # noqa: E402
from openembedding_b200 import checkpoint  # noqa: E402
from openembedding_b200.context import get_context  # noqa: E402
from openembedding_b200.serving.client import ServingClient  # noqa: E402
from openembedding_b200.serving.controller import ModelController  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200000)
ap.add_argument("--dim", type=int, default=64)
ap.add_argument("--nodes", type=int, default=2)
ap.add_argument("--replicas", type=int, default=2)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--compress", default="")
a = ap.parse_args()
oe.flags.device = "cpu"
v = embed.Variable(shape=(a.rows, a.dim), name="v", initializer={"category": "uniform", "minval": -1, "maxval": 1})
ids = torch.arange(a.rows)
v.push_gradients(ids, torch.zeros(a.rows, a.dim))
v.update_weights()
d = tempfile.mkdtemp()
checkpoint.save_model(get_context(), d + "/m", include_optimizer=False)
master = oe.Master()
nodes = [oe.Server(master_endpoint=master.endpoint) for _ in range(a.nodes)]
sign = ModelController(master.endpoint).create_model(d + "/m", replica_num=a.replicas, shard_num=-1)
var = ServingClient(master.endpoint, message_compress=a.compress).find_model_variable(sign, 0)
q = torch.randint(0, a.rows, (a.batch,))
var.pull(q)
t0 = time.time()
for _ in range(a.iters):
    var.pull(q)
dt = (time.time() - t0) / a.iters
print(json.dumps({"rows_per_s": round(a.batch / dt), "ms_per_pull": round(dt * 1e3, 3), "batch": a.batch, "dim": a.dim,
                  "nodes": a.nodes, "replicas": a.replicas, "compress": a.compress}))
for n in nodes:
    n.exit()
