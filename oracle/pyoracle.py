"""ctypes binding of oracle/_build/libcaoracle.so (TEST INFRASTRUCTURE ONLY).

The oracle consumes the very same ``cae_objects`` tables as the engine, but evaluates them with a
CPU restatement of the reference's Go code (oracle/ca_oracle.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcaoracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ca_oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "caengine.h")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B" if force else "-s"])
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        vp, i32, i64p = C.c_void_p, C.c_int32, C.POINTER(C.c_int64)
        l.cao_version.restype = C.c_char_p
        l.cao_feasibility.argtypes = [vp, i32, i32, i32, i32, i32, vp, i64p]
        l.cao_feasibility.restype = i32
        l.cao_estimate.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, i64p]
        l.cao_estimate.restype = i32
        l.cao_estimate_all.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i64p]
        l.cao_estimate_all.restype = i32
        l.cao_estimate_all_li.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
        l.cao_estimate_all_li.restype = i32
        l.cao_pod_score.argtypes = [vp, i32, i32]
        l.cao_pod_score.restype = C.c_double
        l.cao_get_min_limit.argtypes = [C.c_int64, C.c_int64]
        l.cao_get_min_limit.restype = C.c_int64
        l.cao_cluster_capacity_limit.argtypes = [i32, i32, i32]
        l.cao_cluster_capacity_limit.restype = i32
        l.cao_sng_capacity_limit.argtypes = [i32, vp, vp, i32]
        l.cao_sng_capacity_limit.restype = i32
        l.cao_limiter_grants.argtypes = [vp, i32, i32]
        l.cao_limiter_grants.restype = i32
        l.cao_waste_score.argtypes = [vp, i32, i32, vp]
        l.cao_waste_score.restype = C.c_double
        l.cao_expander.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
        l.cao_expander.restype = i32
        l.cao_go_tanh.argtypes = [C.c_double]
        l.cao_go_tanh.restype = C.c_double
        l.cao_price_scores.argtypes = [vp, vp, vp, vp, vp, vp, C.c_double, C.c_int64, vp, vp, vp, vp]
        l.cao_price_scores.restype = i32
        l.cao_expander_ex.argtypes = [i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        l.cao_expander_ex.restype = i32
        l.cao_filter_schedulable.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp]
        l.cao_filter_schedulable.restype = i32
        _lib = l
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def feasibility_dense(enc, p_range: Optional[Tuple[int, int]] = None,
                      t_range: Optional[Tuple[int, int]] = None) -> Tuple[np.ndarray, int]:
    """reasons[t][p] for pending pods x templates (SchedulablePodGroups semantics per pod)."""
    pb, pe = p_range or (0, enc.P)
    tb, te = t_range or (0, enc.T)
    out = np.zeros((te - tb, pe - pb), np.uint8)
    ev = C.c_int64(0)
    rc = lib().cao_feasibility(enc.ptr(), 0, pb, pe, tb, te, _p(out), C.byref(ev))
    assert rc == 0
    return out, ev.value


def feasibility_groups(enc, t_range: Optional[Tuple[int, int]] = None) -> np.ndarray:
    tb, te = t_range or (0, enc.T)
    out = np.zeros((te - tb, enc.E), np.uint8)
    rc = lib().cao_feasibility(enc.ptr(), 1, 0, 0, tb, te, _p(out), None)
    assert rc == 0
    return out


def estimate(enc, tmpl: int, groups: Optional[Sequence[int]] = None, max_nodes: int = 0):
    """One BinpackingNodeEstimator.Estimate on a fresh snapshot fork.
    Returns (node_count, pod_count, sched_count[E], order[list], placements[list of node list idx])."""
    g = np.asarray(list(range(enc.E)) if groups is None else list(groups), np.int32)
    nc, pc = C.c_int32(0), C.c_int32(0)
    sched = np.zeros(enc.E, np.int32)
    order = np.full(enc.E, -1, np.int32)
    plc = np.zeros(max(enc.P, 1), np.int32)
    ev = C.c_int64(0)
    rc = lib().cao_estimate(enc.ptr(), tmpl, _p(g), len(g), max_nodes, C.byref(nc), C.byref(pc),
                            _p(sched), _p(order), _p(plc), C.byref(ev))
    assert rc == 0
    return nc.value, pc.value, sched, [int(x) for x in order if x >= 0], plc[:pc.value].copy()


def estimate_all(enc, max_nodes: Optional[np.ndarray] = None, t_range: Optional[Tuple[int, int]] = None):
    tb, te = t_range or (0, enc.T)
    n = te - tb
    mn = None if max_nodes is None else np.ascontiguousarray(max_nodes, np.int32)
    node_count = np.zeros(n, np.int32)
    pod_count = np.zeros(n, np.int32)
    sched = np.zeros((n, enc.E), np.int32)
    order = np.full((n, enc.E), -1, np.int32)
    ev = C.c_int64(0)
    rc = lib().cao_estimate_all(enc.ptr(), _p(mn), tb, te, _p(node_count), _p(pod_count), _p(sched),
                                _p(order), C.byref(ev))
    assert rc == 0
    return node_count, pod_count, sched, order, ev.value


def estimate_all_li(enc, max_nodes, last_index_in, chain: bool = False):
    """estimate_all with the runner's lastIndex carried in per template (chain: from template to template).
    Returns node_count, pod_count, sched, order, last_index_out."""
    n = enc.T
    mn = None if max_nodes is None else np.ascontiguousarray(max_nodes, np.int32)
    li = np.ascontiguousarray(last_index_in, np.int32)
    node_count, pod_count = np.zeros(n, np.int32), np.zeros(n, np.int32)
    sched, order = np.zeros((n, enc.E), np.int32), np.full((n, enc.E), -1, np.int32)
    lo = np.zeros(n, np.int32)
    rc = lib().cao_estimate_all_li(enc.ptr(), _p(mn), _p(li), int(chain), 0, n, _p(node_count), _p(pod_count), _p(sched), _p(order), _p(lo))
    assert rc == 0
    return node_count, pod_count, sched, order, lo


def pod_score(enc, spec: int, tmpl: int) -> float:
    return lib().cao_pod_score(enc.ptr(), spec, tmpl)


def expander(enc, chain: Sequence[int], node_count, pod_count, sched):
    chain_a = np.asarray(chain, np.int32)
    mask = np.zeros(enc.T, np.uint8)
    waste = np.zeros(enc.T, np.float64)
    rc = lib().cao_expander(enc.ptr(), _p(chain_a), len(chain_a),
                            _p(np.ascontiguousarray(node_count, np.int32)),
                            _p(np.ascontiguousarray(pod_count, np.int32)),
                            _p(np.ascontiguousarray(sched, np.int32)), _p(mask), _p(waste))
    assert rc == 0
    return mask, waste


def go_tanh(x: float) -> float:
    return lib().cao_go_tanh(float(x))


def price_scores(enc, node_price, pod_price, stabilization_price, preferred_cpu_milli=0, unfitness=None, has_gpu=None,
                 exists=None, node_count=None, sched=None, order=None) -> np.ndarray:
    a = lambda x, dt: None if x is None else np.ascontiguousarray(x, dt)
    np_, pp, uf, hg, ex = a(node_price, np.float64), a(pod_price, np.float64), a(unfitness, np.float64), a(has_gpu, np.uint8), a(exists, np.uint8)
    nc, sc, od = a(node_count, np.int32), a(sched, np.int32), a(order, np.int32)
    score = np.zeros(enc.T, np.float64)
    rc = lib().cao_price_scores(enc.ptr(), _p(np_), _p(pp), _p(uf), _p(hg), _p(ex), float(stabilization_price), int(preferred_cpu_milli),
                                _p(nc), _p(sc), _p(od), _p(score))
    assert rc == 0
    return score


def expander_ex(chain: Sequence[int], node_count, pod_count, waste=None, price=None, price_error=None, priority=None) -> np.ndarray:
    a = lambda x, dt: None if x is None else np.ascontiguousarray(x, dt)
    ch, nc, pc = a(chain, np.int32), a(node_count, np.int32), a(pod_count, np.int32)
    w, pr, pe, prio = a(waste, np.float64), a(price, np.float64), a(price_error, np.uint8), a(priority, np.int32)
    mask = np.zeros(len(nc), np.uint8)
    rc = lib().cao_expander_ex(len(nc), _p(ch), len(ch), _p(nc), _p(pc), _p(w), _p(pr), _p(pe), _p(prio), _p(mask))
    assert rc == 0
    return mask


def get_min_limit(base: int, target: int) -> int:
    return lib().cao_get_min_limit(base, target)


def limiter_grants(limits: Sequence[int], asks: int) -> int:
    a = np.asarray(limits, np.int32)
    return lib().cao_limiter_grants(_p(a), len(a), asks)


def cluster_capacity_limit(has_ctx: bool, max_limit: int, current: int) -> int:
    return lib().cao_cluster_capacity_limit(int(has_ctx), max_limit, current)


def sng_capacity_limit(has_ctx: bool, max_sizes: Sequence[int], target_sizes: Sequence[int]) -> int:
    a, b = np.asarray(max_sizes, np.int32), np.asarray(target_sizes, np.int32)
    return lib().cao_sng_capacity_limit(int(has_ctx), _p(a), _p(b), len(a))


def filter_schedulable(enc, pod_order: Sequence[int], hint_node=None, sim_class=None, class_ctrl=None, node_ok=None,
                       last_index: int = 0, break_on_failure: bool = False):
    """filterOutSchedulableByPacking / HintingSimulator.TrySchedulePods on the cluster snapshot.
    Returns (assigned[num_pending] node index or -1, lastIndex afterwards, overflowing controller count)."""
    order = np.ascontiguousarray(pod_order, np.int32)
    hn = None if hint_node is None else np.ascontiguousarray(hint_node, np.int32)
    sc = None if sim_class is None else np.ascontiguousarray(sim_class, np.int32)
    cc = None if class_ctrl is None else np.ascontiguousarray(class_ctrl, np.int32)
    ok = None if node_ok is None else np.ascontiguousarray(node_ok, np.uint8)
    assigned = np.full(enc.P, -1, np.int32)
    li = np.zeros(1, np.int32)
    ov = np.zeros(1, np.int32)
    rc = lib().cao_filter_schedulable(enc.ptr(), _p(order), len(order), _p(hn), _p(sc), _p(cc), _p(ok), int(last_index),
                                      int(break_on_failure), _p(assigned), _p(li), _p(ov))
    assert rc == 0
    return assigned, int(li[0]), int(ov[0])
