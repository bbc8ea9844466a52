"""Deterministic synthetic snapshots for the BASELINE.json configs.

Generator = splitmix64 streams, seed ``0xCA5CA1E0 + config_index``.  Tables are built at the
integer level (``TableBuilder``) — there are no strings to intern in a synthetic snapshot, the ids
stand for them (key 0 = kubernetes.io/hostname, 1 = zone, 2 = pool, 3 = instance-type, 4 = app,
5 = tier, 6.. = taint keys).

Distributions: pods in E = P/100 equivalence groups with Zipf(1.1) replica counts; cpu in
{50,100,250,500,1000,2000,4000} m (20/25/20/15/10/7/3 %), mem = cpu x {1,2,4,8} MiB, 10 % of the
groups ask for nvidia.com/gpu in {1,2,4,8}; 32 namespaces; labels app=<group>, tier in 4.
Templates: vCPU in {2,..,96}, allocatable = capacity - reserved, 110 pods, 2 DaemonSet pods, 15 % GPU
templates, labels hostname/zone(16)/pool(8)/instance-type.  C2+: 16 taints (8 keys x 2 values), 0-2
NoSchedule taints per template, tolerations per group (10 % wildcard Exists), 20 % of the groups
carry a 1-2 key nodeSelector.  C3+: 30 % of the groups have one DoNotSchedule spread constraint and
the snapshot holds `cluster_nodes` nodes x 30 resident pods.  C4+: 10 % self anti-affinity on
hostname, 5 % required affinity to another group on zone.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .encode import EncodedObjects, TableBuilder

MiB = 1 << 20
GiB = 1 << 30
K_HOST, K_ZONE, K_POOL, K_ITYPE, K_APP, K_TIER, K_TAINT0 = 0, 1, 2, 3, 4, 5, 6
RES_GPU = 3


@dataclass
class Config:
    index: int
    pods: int
    templates: int
    taints: bool = False
    spread: bool = False
    affinity: bool = False
    cluster_nodes: int = 0
    pods_per_node: int = 30
    name: str = ""


CONFIGS = {
    1: Config(1, 1_000, 50, name="C1 1000 pods x 50 templates, NodeResourcesFit only"),
    2: Config(2, 100_000, 1_000, taints=True, name="C2 100k pods x 1000 templates, resources + taints/tolerations"),
    3: Config(3, 100_000, 5_000, taints=True, spread=True, cluster_nodes=2000,
              name="C3 100k pods x 5000 templates, + PodTopologySpread"),
    4: Config(4, 500_000, 5_000, taints=True, spread=True, affinity=True, cluster_nodes=2000,
              name="C4 500k pods x 5000 templates, + InterPodAffinity"),
    5: Config(5, 1_000_000, 10_000, taints=True, spread=True, affinity=True, cluster_nodes=2000,
              name="C5 1M pods x 10000 templates, full predicate set"),
}


class SplitMix64:
    """Vectorised splitmix64 stream."""

    def __init__(self, seed: int) -> None:
        self.state = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)

    def next(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = z[-1] if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n: int) -> np.ndarray:
        return (self.next(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)

    def randint(self, n: int, hi: int) -> np.ndarray:
        return (self.next(n) % np.uint64(hi)).astype(np.int64)

    def choice(self, n: int, values, weights) -> np.ndarray:
        cdf = np.cumsum(np.asarray(weights, dtype=np.float64))
        cdf /= cdf[-1]
        return np.asarray(values)[np.searchsorted(cdf, self.uniform(n), side="right").clip(0, len(values) - 1)]


RESERVED = {2: 70, 4: 80, 8: 90, 16: 110, 32: 150, 64: 230, 96: 310}


def generate(config: int = 2, pods: Optional[int] = None, templates: Optional[int] = None,
             cluster_nodes: Optional[int] = None, taints: Optional[bool] = None,
             spread: Optional[bool] = None, affinity: Optional[bool] = None,
             seed: Optional[int] = None) -> EncodedObjects:
    cfg = CONFIGS[config]
    P = cfg.pods if pods is None else pods
    T = cfg.templates if templates is None else templates
    NC = cfg.cluster_nodes if cluster_nodes is None else cluster_nodes
    use_taints = cfg.taints if taints is None else taints
    use_spread = cfg.spread if spread is None else spread
    use_aff = cfg.affinity if affinity is None else affinity
    rng = SplitMix64((0xCA5CA1E0 + cfg.index) if seed is None else seed)
    b = TableBuilder(num_res=4)
    b.hostname_key = K_HOST
    b.unschedulable_taint_key = -1
    for ns in range(32):
        b.declare_namespace(ns, 0, False)

    # value id spaces (all label values are opaque ids; none parses as an int)
    V_ZONE0, V_POOL0, V_ITYPE0, V_TIER0, V_TAINTV0 = 0, 16, 24, 48, 52
    V_APP0 = 64
    E = max(1, P // 100)
    V_HOST0 = V_APP0 + E
    b.declare_value(V_HOST0 + T + NC + 1, None)

    # ---- groups ------------------------------------------------------------------------------
    w = 1.0 / np.arange(1, E + 1, dtype=np.float64) ** 1.1
    counts = np.maximum(1, np.floor(P * w / w.sum()).astype(np.int64))
    diff = P - int(counts.sum())
    i = 0
    while diff != 0:  # hand the rounding remainder to the head groups
        step = 1 if diff > 0 else -1
        if counts[i % E] + step >= 1:
            counts[i % E] += step
            diff -= step
        i += 1
    perm = rng.next(E).argsort()  # which group is big is random, not tied to its id
    counts = counts[perm]
    cpu = rng.choice(E, [50, 100, 250, 500, 1000, 2000, 4000], [20, 25, 20, 15, 10, 7, 3]).astype(np.int64)
    mem = cpu * rng.choice(E, [1, 2, 4, 8], [1, 1, 1, 1]).astype(np.int64) * MiB
    gpu = np.where(rng.uniform(E) < 0.10, rng.choice(E, [1, 2, 4, 8], [1, 1, 1, 1]), 0).astype(np.int64)
    ns_of = rng.randint(E, 32)
    tier = rng.randint(E, 4)
    u_tol = rng.uniform(E)
    tol_bits = rng.next(E)
    u_sel = rng.uniform(E)
    sel_pool = rng.randint(E, 8)
    sel_zone = rng.randint(E, 16)
    u_pts = rng.uniform(E)
    pts_kind = rng.randint(E, 2)
    pts_skew_h = rng.choice(E, [1, 2], [1, 1])
    pts_skew_z = rng.choice(E, [1, 2, 5], [1, 1, 1])
    pts_mind = rng.choice(E, [1, 3], [1, 1])
    u_aff = rng.uniform(E)
    aff_other = rng.randint(E, max(E, 1))

    taint_ids = [(K_TAINT0 + k, V_TAINTV0 + v) for k in range(8) for v in range(2)]
    group_spec: List[int] = []
    for g in range(E):
        ls = b.labelset([(K_APP, V_APP0 + g), (K_TIER, V_TIER0 + int(tier[g]))])
        tols = []
        if use_taints:
            if u_tol[g] < 0.10:
                tols.append((-1, 1, -1, 0))  # wildcard: empty key + Exists tolerates everything
            else:
                for i, (k, v) in enumerate(taint_ids):
                    if (int(tol_bits[g]) >> i) & 1:
                        tols.append((k, 0, v, 1))  # key=value:NoSchedule
        naff = -1
        if use_taints and u_sel[g] < 0.20:
            reqs = [(K_POOL, 0, (V_POOL0 + int(sel_pool[g]),))]
            if u_sel[g] < 0.07:
                reqs.append((K_ZONE, 0, (V_ZONE0 + int(sel_zone[g]),)))
            naff = b.node_affinity(b.selector(reqs), False, [])
        pts = 0
        anti = 0
        aff = 0
        own_sel = b.selector([(K_APP, 0, (V_APP0 + g,))]) if (use_spread or use_aff) else 0
        if use_spread and u_pts[g] < 0.30:
            if pts_kind[g] == 0:
                pts = b.pts_list([(int(pts_skew_h[g]), K_HOST, own_sel, int(pts_mind[g]), 1, 0)])
            else:
                pts = b.pts_list([(int(pts_skew_z[g]), K_ZONE, own_sel, int(pts_mind[g]), 1, 0)])
        if use_aff:
            nothing = b.nothing_selector()
            if u_aff[g] < 0.10:
                anti = b.affinity_list([(own_sel, K_HOST, (int(ns_of[g]),), nothing)])
            elif u_aff[g] < 0.15:
                og = int(aff_other[g])
                osel = b.selector([(K_APP, 0, (V_APP0 + og,))])
                aff = b.affinity_list([(osel, K_ZONE, (int(ns_of[og]),), nothing)])
        req = [int(cpu[g]), int(mem[g]), 0, int(gpu[g])]
        group_spec.append(b.podspec(int(ns_of[g]), ls, req, b.toleration_list(tols), naff, -1, 0, pts, aff, anti))

    # ---- DaemonSet pods (kube-system = namespace 31, no labels anyone selects) --------------------
    ds_ls = b.labelset([(K_TIER, V_TIER0 + 3)])
    ds_tol = b.toleration_list([(-1, 1, -1, 0)])
    ds_spec = [b.podspec(31, ds_ls, [100, 200 * MiB, 0, 0], ds_tol), b.podspec(31, ds_ls, [100, 200 * MiB, 0, 0], ds_tol)]

    # ---- nodes ----------------------------------------------------------------------------------
    def node_shape(n: int, r: SplitMix64):
        vcpu = r.choice(n, [2, 4, 8, 16, 32, 64, 96], [1] * 7).astype(np.int64)
        mem_per = r.choice(n, [2, 4, 8], [1, 1, 1]).astype(np.int64)
        is_gpu = r.uniform(n) < 0.15
        ngpu = np.where(is_gpu, r.choice(n, [1, 4, 8], [1, 1, 1]), 0).astype(np.int64)
        return vcpu, mem_per, ngpu, r.randint(n, 16), r.randint(n, 8), r.uniform(n), r.next(n)

    def add_nodes(n: int, first_host: int, is_template: bool, resident: Optional[np.ndarray]):
        vcpu, mem_per, ngpu, zone, pool, u_t, tbits = node_shape(n, rng)
        for i in range(n):
            v = int(vcpu[i])
            cap_cpu = v * 1000
            cap_mem = v * int(mem_per[i]) * GiB
            alloc_cpu = cap_cpu - RESERVED[v] - (i % 7)  # a few milli-cores of jitter: distinct allocatable
            alloc_mem = cap_mem - cap_mem // 20 - (i % 11) * MiB
            ls = b.labelset([(K_HOST, V_HOST0 + first_host + i), (K_ZONE, V_ZONE0 + int(zone[i])),
                             (K_POOL, V_POOL0 + int(pool[i])), (K_ITYPE, V_ITYPE0 + (v % 24))])
            tl = []
            if use_taints:
                nt = 0 if u_t[i] < 0.5 else (1 if u_t[i] < 0.85 else 2)
                for j in range(nt):
                    k, val = taint_ids[(int(tbits[i]) >> (8 * j)) % 16]
                    tl.append((k, val, 1))
            pods = list(ds_spec)
            if resident is not None:
                pods += [group_spec[int(x)] for x in resident[i]]
            args = dict(name=first_host + i, labelset=ls, taint_list=b.taint_list(tl), unschedulable=False,
                        alloc=[alloc_cpu, alloc_mem, 0, int(ngpu[i])], allowed_pods=110, cap_cpu=cap_cpu,
                        cap_mem=cap_mem, has_alloc_cpu=True, has_alloc_mem=True, pod_specs=pods)
            (b.template if is_template else b.cluster_node)(**args)

    if NC:
        resident = rng.randint(NC * cfg.pods_per_node, E).reshape(NC, cfg.pods_per_node)
        add_nodes(NC, 0, False, resident)
    add_nodes(T, NC, True, None)

    # ---- pending pods, group-major ------------------------------------------------------------------
    for g in range(E):
        b.group(np.full(int(counts[g]), group_spec[g], np.int32))
    return b.finish()
