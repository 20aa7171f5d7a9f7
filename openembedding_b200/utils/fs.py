"""URI-addressed file access for checkpoints.

Reference: pico-core ``URIConfig`` (``scheme://path?k=v`` with per-URI parameters),
``FileSystem`` / ``ShellUtility`` (local files or ``hdfs dfs -cat/-put`` pipes,
pico-ps/pico-core/src/common/FileSystem.h, ShellUtility.h) and the ``mem://null/`` sink of
the dump operator (openembedding/server/EmbeddingDumpOperator.cpp).

Local paths are used directly (the native shard writer streams to them). ``hdfs://`` models are
staged: ``stage_out`` hands the writer a local scratch directory and ``commit`` uploads it with
``hdfs dfs -put``; ``stage_in`` downloads with ``hdfs dfs -get``. ``mem://null/`` discards.
"""
import os
import shutil
import subprocess
import tempfile
from urllib.parse import parse_qsl, urlparse


class URIConfig:
    def __init__(self, uri):
        self.uri = str(uri)
        u = urlparse(self.uri)
        self.scheme = u.scheme if u.scheme and len(u.scheme) > 1 else "file"
        if self.scheme == "file":
            path = self.uri[len("file://"):] if self.uri.startswith("file://") else self.uri
            self.path, _, q = path.partition("?")
        else:
            self.path, q = self.uri.split("?", 1)[0], u.query
        self.params = dict(parse_qsl(q))

    @property
    def is_local(self):
        return self.scheme == "file"

    @property
    def is_null(self):
        return self.scheme == "mem" and self.uri.startswith("mem://null")

    def __str__(self):
        q = "&".join("%s=%s" % kv for kv in self.params.items())
        return self.path + ("?" + q if q else "")


def _hdfs(*args):
    exe = os.environ.get("HADOOP_BIN", "hdfs")
    return subprocess.run([exe, "dfs"] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE)


class Staging:
    """context manager: ``with Staging(uri, "w") as local_dir: ...write...`` (commit on success)"""

    def __init__(self, uri, mode="r"):
        self.cfg, self.mode, self.tmp = URIConfig(uri), mode, None

    def __enter__(self):
        c = self.cfg
        if c.is_local:
            if self.mode == "w":
                os.makedirs(c.path, exist_ok=True)
            return c.path
        self.tmp = tempfile.mkdtemp(prefix="exb_stage_")
        if c.is_null:
            return self.tmp
        if c.scheme == "hdfs":
            if self.mode == "r":
                r = _hdfs("-get", c.path + "/*", self.tmp)
                if r.returncode != 0:
                    raise IOError("hdfs get failed: " + r.stderr.decode(errors="replace")[-500:])
            return self.tmp
        raise ValueError("unsupported uri scheme: " + c.scheme)

    def __exit__(self, et, ev, tb):
        c = self.cfg
        try:
            if et is None and self.tmp and self.mode == "w" and c.scheme == "hdfs":
                _hdfs("-mkdir", "-p", c.path)
                r = _hdfs("-put", "-f", *(os.path.join(self.tmp, f) for f in os.listdir(self.tmp)), c.path)
                if r.returncode != 0:
                    raise IOError("hdfs put failed: " + r.stderr.decode(errors="replace")[-500:])
        finally:
            if self.tmp:
                shutil.rmtree(self.tmp, ignore_errors=True)
        return False


def exists(uri):
    c = URIConfig(uri)
    if c.is_local:
        return os.path.exists(c.path)
    if c.scheme == "hdfs":
        return _hdfs("-test", "-e", c.path).returncode == 0
    return False
