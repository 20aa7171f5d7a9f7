"""Scoped timers and accumulators.

Reference: ``VTIMER(level, component, name, unit)`` scoped timers feeding distributed
``Accumulator<TimerAggregator>`` tables that rank 0 prints every ``server.report_interval``
seconds (pico-core accumulator/AutoTimer.h:214-243, openembedding/client/WorkerContext.cpp:
140-163), plus the ``pull_indices`` / ``pull_unique`` counters
(openembedding/server/EmbeddingPullOperator.cpp:208-247).

Here: host-side ``vtimer`` context managers (wall clock or CUDA events), named counters, and
the device-side phase clock of the fused push+update kernel (``%globaltimer`` stamps read back
through ``CudaEngine.status()``). ``report()`` renders the same kind of table; ``reduce()``
sums it over ranks.
"""
import contextlib
import threading
import time

import torch

_lock = threading.Lock()
_timers = {}     # name -> [count, total_ms, max_ms]
_counters = {}   # name -> value
VTIMER_LEVEL = 1
enabled = False  # reference: active only when server.report_interval > 0


def enable(on=True):
    global enabled
    enabled = bool(on)


def add_time(name, ms):
    with _lock:
        t = _timers.setdefault(name, [0, 0.0, 0.0])
        t[0] += 1
        t[1] += ms
        t[2] = max(t[2], ms)


def add_count(name, value=1):
    with _lock:
        _counters[name] = _counters.get(name, 0) + value


@contextlib.contextmanager
def vtimer(level, component, name, cuda=False):
    """with vtimer(1, "client", "pull"): ...   (no-op unless enabled and level <= VTIMER_LEVEL)"""
    if not enabled or level > VTIMER_LEVEL:
        yield
        return
    key = "%s.%s" % (component, name)
    if cuda and torch.cuda.is_available():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        e1.synchronize()
        add_time(key, e0.elapsed_time(e1))
    else:
        t0 = time.perf_counter()
        yield
        add_time(key, (time.perf_counter() - t0) * 1e3)


def snapshot(ctx=None):
    """dict of all timers/counters; with a context also the engine's device-side counters"""
    with _lock:
        out = {"timers": {k: {"count": v[0], "total_ms": v[1], "max_ms": v[2],
                              "avg_ms": v[1] / v[0] if v[0] else 0.0} for k, v in _timers.items()},
               "counters": dict(_counters)}
    if ctx is not None and getattr(ctx.backend, "name", "") == "cuda":
        code, st = ctx.backend.engine.status()
        out["counters"].update({"pull_indices": st["pull_indices"], "push_indices": st["push_indices"],
                                "update_unique": st["update_unique"]})
        out["push_update_phases_us"] = st.get("last_push_update_us")
    elif ctx is not None:
        out["counters"].update(ctx.backend.counters)
    return out


def reduce(ctx, snap=None):
    """sum counters / timer totals over all ranks (rank 0 hosts the AccumulatorServer in the reference)"""
    snap = snap or snapshot(ctx)
    if not ctx.dist_on:
        return snap
    import torch.distributed as dist
    objs = [None] * ctx.world
    dist.all_gather_object(objs, snap, group=ctx.group)
    out = {"timers": {}, "counters": {}}
    for s in objs:
        for k, v in s["timers"].items():
            t = out["timers"].setdefault(k, {"count": 0, "total_ms": 0.0, "max_ms": 0.0})
            t["count"] += v["count"]; t["total_ms"] += v["total_ms"]; t["max_ms"] = max(t["max_ms"], v["max_ms"])
        for k, v in s["counters"].items():
            out["counters"][k] = out["counters"].get(k, 0) + v
    for t in out["timers"].values():
        t["avg_ms"] = t["total_ms"] / t["count"] if t["count"] else 0.0
    return out


def report(ctx=None, file=None):
    s = snapshot(ctx)
    lines = ["%-40s %10s %12s %12s %12s" % ("timer", "count", "total_ms", "avg_ms", "max_ms")]
    for k in sorted(s["timers"]):
        v = s["timers"][k]
        lines.append("%-40s %10d %12.3f %12.4f %12.4f" % (k, v["count"], v["total_ms"], v["avg_ms"], v["max_ms"]))
    for k in sorted(s["counters"]):
        lines.append("%-40s %10d" % (k, s["counters"][k]))
    pi, pu = s["counters"].get("pull_indices", 0), s["counters"].get("pull_unique", 0)
    if pi and pu:
        lines.append("%-40s %10.4f" % ("pull_unique/pull_indices", pu / pi))
    text = "\n".join(lines)
    print(text, file=file)
    return text


def reset():
    with _lock:
        _timers.clear()
        _counters.clear()


class Monitor:
    """Periodic reporter thread (pico-core common/Monitor.h): prints the table every `interval` s."""

    def __init__(self, ctx, interval):
        self.ctx, self.interval = ctx, float(interval)
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)

    def start(self):
        enable(True)
        self._th.start()
        return self

    def _run(self):
        while not self._stop.wait(self.interval):
            if self.ctx.rank == 0:
                report(None)

    def stop(self):
        self._stop.set()


# ---- NVTX ranges ("CUDA events per phase ... NVTX ranges"). EXB_NVTX=1 turns them on; off they cost nothing.
import os as _os

nvtx_enabled = _os.environ.get("EXB_NVTX", "0") == "1"


class nvtx_range:
    """``with nvtx_range("pull"):`` -- a named range on the timeline of nsys / ncu when EXB_NVTX=1"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx_enabled:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *a):
        if nvtx_enabled:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


def nvtx_mark(name):
    if nvtx_enabled:
        import torch
        torch.cuda.nvtx.mark(name)
