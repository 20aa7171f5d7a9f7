"""Serve an exported model: master + N serving nodes + REST controller, then pull through the
client -- counterpart of examples/tensorflow_serving_{restful,client}.py + run/*.sh.

    python examples/serving_restful.py --model /tmp/exb_hook_example/openembedding --nodes 3
"""
import argparse
import json
import os
import sys
import threading
import urllib.request

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openembedding_b200 as oe  # noqa: E402
from openembedding_b200.serving.client import ServingClient  # noqa: E402
from openembedding_b200.serving.controller import serve  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", required=True, help="directory written by save_server_model / model.save(...)/openembedding")
ap.add_argument("--nodes", type=int, default=3)
ap.add_argument("--replicas", type=int, default=2)
ap.add_argument("--port", type=int, default=8010)
a = ap.parse_args()

master = oe.Master()
nodes = [oe.Server(master_endpoint=master.endpoint) for _ in range(a.nodes)]
httpd, ctl = serve(master.endpoint, port=a.port, bind_ip="127.0.0.1")
threading.Thread(target=httpd.serve_forever, daemon=True).start()
base = "http://127.0.0.1:%d" % httpd.server_address[1]
req = urllib.request.Request(base + "/models", method="POST", headers={"Content-Type": "application/json"},
                             data=json.dumps({"model_uri": a.model, "replica_num": a.replicas, "shard_num": -1}).encode())
sign = json.loads(urllib.request.urlopen(req).read())["model_sign"]
print("model", sign, json.loads(urllib.request.urlopen(base + "/models/" + sign).read())["model_status"])
cli = ServingClient(master.endpoint)
print(cli.find_model_variable(sign, 0).pull(torch.arange(5)))
for n in nodes:
    n.exit()
httpd.shutdown()
