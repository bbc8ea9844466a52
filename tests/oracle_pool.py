"""Worker pool around the CPU oracle (TEST INFRASTRUCTURE): every worker regenerates the synthetic
snapshot (deterministic, < 1 s) and answers oracle queries on template / pod slices, so the full-size
BASELINE configurations can be checked in about a minute on the GPU box's host cores."""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_W = {}


def cores(limit: int = 32) -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:   # cgroup v2 CPU quota, if any
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        pass
    return max(1, min(n, limit))


def _init(config, kwargs):
    from kubernetes_autoscaler_b200 import synth
    from oracle import pyoracle
    _W["enc"] = synth.generate(config, **kwargs)
    _W["oracle"] = pyoracle
    pyoracle.lib()


def _dense(job):
    p_range, t_range = job
    return t_range, p_range, _W["oracle"].feasibility_dense(_W["enc"], p_range=p_range, t_range=t_range)[0]


def _groups(t):
    return t, _W["oracle"].feasibility_groups(_W["enc"], t_range=(t, t + 1))[0]


def _estimate(job):
    import numpy as np
    t, cap = job
    enc = _W["enc"]
    nc, pc, sched, order, _ = _W["oracle"].estimate_all(enc, np.full(enc.T, cap, np.int32), t_range=(t, t + 1))
    return t, int(nc[0]), int(pc[0]), sched[0], order[0]


class OraclePool:
    def __init__(self, config: int, procs: int = 0, **kwargs):
        self.procs = procs or cores()
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_init, initargs=(config, kwargs))

    def dense(self, jobs):
        """jobs: [((p_begin, p_end), (t_begin, t_end))] -> [(t_range, p_range, reasons[t][p])]"""
        return self.pool.map(_dense, jobs, chunksize=1)

    def groups(self, templates):
        return dict(self.pool.map(_groups, list(templates), chunksize=1))

    def estimate(self, templates, cap):
        return {r[0]: r[1:] for r in self.pool.map(_estimate, [(t, cap) for t in templates], chunksize=1)}

    def close(self):
        self.pool.terminate()
        self.pool.join()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
