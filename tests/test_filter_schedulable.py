"""filter-out-schedulable / HintingSimulator.

CPU part: the oracle's restatement against the reference's own test tables
(simulator/scheduling/hinting_simulator_test.go:32-182,184-260; core/podlistprocessor/filter_out_schedulable_test.go:35-210).
GPU part: the engine through the C ABI against the oracle, bit-exact (assigned node per pod, lastIndex,
overflowing controllers)."""
import numpy as np
import pytest

from kubernetes_autoscaler_b200 import podlistprocessor as plp
from kubernetes_autoscaler_b200.objects import (BuildTestNode, BuildTestPod, LabelSelector, NodeInfo, PodAffinityTerm, Taint,
                                                TopologySpreadConstraint, WithLabels, WithPodAffinity, WithPodAntiAffinity, makeNode)


class OracleSimulator(plp.HintingSimulator):
    """Same host logic, placement loop on the CPU oracle."""

    def _run(self, x, breakOnFailure):
        from oracle import pyoracle
        return pyoracle.filter_schedulable(x.enc, x.order, x.hint, x.sim_class, x.class_ctrl, x.node_ok, self.last_index, breakOnFailure)


def _ready_node(name, cpu, mem):
    return BuildTestNode(name, cpu, mem)


def _scheduled(name, cpu, mem):
    return BuildTestPod(name, cpu, mem)


def _two_nodes():
    return [NodeInfo(_ready_node("n1", 1000, 2000000), [_scheduled("p1", 300, 500000)]), NodeInfo(_ready_node("n2", 1000, 2000000))]


TRY_CASES = [
    # (new pods (name, cpu, mem), hints {name: node}, acceptable node or None, expected [(pod, node)])
    ("two new pods, two nodes", [("p2", 800), ("p3", 500)], {}, None, [("p2", "n2"), ("p3", "n1")]),
    ("hinted Node no longer in the cluster", [("p2", 800), ("p3", 500)], {"p2": "non-existing-node"}, None, [("p2", "n2"), ("p3", "n1")]),
    ("three new pods, two nodes, no fit", [("p2", 800), ("p3", 500), ("p4", 700)], {}, None, [("p2", "n2"), ("p3", "n1")]),
    ("no new pods, two nodes", [], {}, None, []),
    ("two nodes, but only one acceptable", [("p2", 500), ("p3", 500)], {}, "n2", [("p2", "n2"), ("p3", "n2")]),
    ("two nodes, but only one acceptable, no fit", [("p2", 500), ("p3", 500)], {}, "n1", [("p2", "n1")]),
]


def _run_try_case(sim, case):
    _, new, hints, only, want = case
    pods = [BuildTestPod(n, c, 500000) for n, c in new]
    for p in pods:
        if p.name in hints:
            sim.hints.Set(plp.HintKeyFromPod(p), hints[p.name])
    ok = plp.ScheduleAnywhere if only is None else (lambda ni: ni.node.name == only)
    statuses, _ = sim.TrySchedulePods(_two_nodes(), pods, ok, False)
    assert [(s.pod.name, s.node_name) for s in statuses] == want
    sim.DropOldHints()
    for pod_name, node in want:   # new hints match the nodes actually used
        assert sim.hints.Get(("default", pod_name)) == node


@pytest.mark.parametrize("case", TRY_CASES, ids=[c[0] for c in TRY_CASES])
def test_oracle_try_schedule_pods_kat(case):
    _run_try_case(OracleSimulator(), case)


HINT_CASES = [
    {"p1": "n2"},
    {"p1": "n2", "p2": "n2", "p3": "n2"},
    {"p1": "n1", "p2": "n2", "p3": "n3"},
    {"p1": "n1", "p2": "n1", "p3": "n1", "p4": "n2", "p5": "n2", "p6": "n2", "p7": "n3", "p8": "n3", "p9": "n3"},
]


def _run_hint_case(sim, pod_nodes):
    cluster = [NodeInfo(_ready_node(n, 9999, 9999)) for n in ("n1", "n2", "n3")]
    pods = [BuildTestPod(p, 1, 1) for p in pod_nodes]
    for p in pods:
        sim.hints.Set(plp.HintKeyFromPod(p), pod_nodes[p.name])
    statuses, _ = sim.TrySchedulePods(cluster, pods)
    assert [(s.pod.name, s.node_name) for s in statuses] == list(pod_nodes.items())


@pytest.mark.parametrize("pod_nodes", HINT_CASES)
def test_oracle_pod_schedules_on_hinted_node_kat(pod_nodes):
    _run_hint_case(OracleSimulator(), pod_nodes)


def _prio(name, cpu, mem, prio):
    p = BuildTestPod(name, cpu, mem)
    p.priority = prio
    return p


FILTER_CASES = [
    # (resident pods on the single 2000m node, candidates, expected scheduled names, expected unscheduled names, node filter)
    ("single empty node, no pods", [], [], [], [], True),
    ("single empty node, single schedulable pod", [], [("pod", 500, 0)], ["pod"], [], True),
    ("single empty node, many schedulable pods", [], [("pod1", 200, 0), ("pod2", 500, 0), ("pod3", 800, 0)], ["pod1", "pod2", "pod3"], [], True),
    ("single empty node, single unschedulable pod", [], [("pod1", 3000, 0)], [], ["pod1"], True),
    ("single empty node, various pods", [], [("pod1", 200, 0), ("pod2", 500, 0), ("pod3", 1800, 0)], ["pod1", "pod2"], ["pod3"], True),
    ("single empty node, some priority pods", [], [("pod1", 200, 0), ("pod2", 500, 10), ("pod3", 1800, 20)], ["pod3", "pod1"], ["pod2"], True),
    ("non-empty node with a single pods scheduled", [500], [("pod2", 1000, 0), ("pod3", 300, 0), ("pod4", 300, 0)], ["pod2", "pod3"], ["pod4"], True),
    ("non-empty node with many pods scheduled", [500, 1000], [("pod3", 1000, 0), ("pod4", 300, 0), ("pod5", 300, 0)], ["pod4"], ["pod3", "pod5"], True),
    ("node should not be considered", [], [("pod1", 200, 0), ("pod2", 500, 0), ("pod3", 1800, 0)], [], ["pod1", "pod2", "pod3"], False),
]


def _run_filter_case(make_processor, case):
    _, resident, cands, want_sched, want_unsched, all_nodes = case
    node = NodeInfo(BuildTestNode("node", 2000, 100), [BuildTestPod("r%d" % i, c, 10) for i, c in enumerate(resident)])
    pods = [_prio(n, c, 10, pr) for n, c, pr in cands]
    proc = make_processor((lambda ni: True) if all_nodes else (lambda ni: False))
    left = proc.Process([node], pods)
    assert sorted(p.name for p in left) == sorted(want_unsched)
    assert sorted(p.name for p in pods if p not in left) == sorted(want_sched)


def _oracle_processor(node_filter):
    proc = plp.FilterOutSchedulablePodListProcessor(node_filter)
    proc.schedulingSimulator = OracleSimulator()
    return proc


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_oracle_filter_out_schedulable_kat(case):
    _run_filter_case(_oracle_processor, case)


# ---- semantics the reference tests do not pin: similar pods, overflow, unschedulable nodes, spread / affinity --------
def _owned(name, cpu, uid, labels=None, kind="ReplicaSet"):
    p = BuildTestPod(name, cpu, 10)
    p.owner_uid, p.owner_kind = uid, kind
    if labels:
        p.labels = dict(labels)
    return p


def _scenarios():
    out = []
    # 1. the similar-pods shortcut CHANGES the result: r2 would fit after "friend" lands (affinity), but a similar pod failed first
    aff = PodAffinityTerm(LabelSelector({"app": "friend"}), "kubernetes.io/hostname")
    cluster = [NodeInfo(makeNode(4000, 4000, 10, "n1", "z1")), NodeInfo(makeNode(4000, 4000, 10, "n2", "z1"))]
    r1, r2 = _owned("r1", 100, "rs-a"), _owned("r2", 100, "rs-a")
    for r in (r1, r2):
        WithPodAffinity(aff)(r)
    friend = BuildTestPod("friend", 100, 10, WithLabels({"app": "friend"}))
    lone = BuildTestPod("lone", 100, 10, WithPodAffinity(aff))   # no controller: evaluated on its own, fits next to friend
    out.append(("similar shortcut", cluster, [r1, friend, r2, lone]))
    # 2. more than 10 distinct unschedulable specs of one controller: overflowing, every pod evaluated
    cluster = [NodeInfo(BuildTestNode("n1", 1000, 1 << 30))]
    pods = [_owned("big%d" % i, 2000 + i, "rs-over") for i in range(13)] + [_owned("big%d-b" % i, 2000 + i, "rs-over") for i in range(13)]
    out.append(("overflowing controller", cluster, pods))
    # 3. unschedulable / tainted nodes are skipped, lastIndex wraps, DaemonSet pods never enter the shortcut
    n_bad = BuildTestNode("n-unsched", 4000, 1 << 30)
    n_bad.unschedulable = True
    n_taint = BuildTestNode("n-taint", 4000, 1 << 30)
    n_taint.taints = [Taint("k", "v", "NoSchedule")]
    cluster = [NodeInfo(n_bad), NodeInfo(BuildTestNode("n1", 1000, 1 << 30)), NodeInfo(n_taint), NodeInfo(BuildTestNode("n2", 1000, 1 << 30))]
    pods = [_owned("a%d" % i, 300, "rs-b") for i in range(8)] + [_owned("ds%d" % i, 900, "ds-1", kind="DaemonSet") for i in range(3)]
    out.append(("unschedulable nodes", cluster, pods))
    # 4. hostname spread + anti-affinity over existing nodes with resident pods
    cluster = [NodeInfo(makeNode(4000, 4000, 10, "n%d" % i, "z%d" % (i % 2)),
                        [BuildTestPod("res%d" % i, 500, 10, WithLabels({"app": "web"}))] if i % 3 == 0 else []) for i in range(7)]
    spread = TopologySpreadConstraint(1, "kubernetes.io/hostname", LabelSelector({"app": "web"}))
    web = []
    for i in range(9):
        p = _owned("web%d" % i, 400, "rs-web", {"app": "web"})
        p.topology_spread = [spread]
        web.append(p)
    anti = [BuildTestPod("solo%d" % i, 200, 10, WithLabels({"app": "solo"}),
                         WithPodAntiAffinity(PodAffinityTerm(LabelSelector({"app": "solo"}), "topology.kubernetes.io/zone"))) for i in range(3)]
    out.append(("spread and anti-affinity", cluster, web[:4] + anti + web[4:]))
    return out


SCENARIOS = _scenarios()


def test_oracle_similar_pods_semantics():
    sim = OracleSimulator()
    name, cluster, pods = SCENARIOS[0]
    st, over = sim.TrySchedulePods(cluster, pods)
    # r1 fails (no friend yet) -> rs-a/spec marked; friend lands; r2 is skipped although it would fit; lone fits
    assert [(s.pod.name) for s in st] == ["friend", "lone"] and over == 0
    sim = OracleSimulator()
    st, over = sim.TrySchedulePods(SCENARIOS[1][1], SCENARIOS[1][2])
    assert st == [] and over == 1
    sim = OracleSimulator()
    st, _ = sim.TrySchedulePods(SCENARIOS[2][1], SCENARIOS[2][2])
    assert {s.node_name for s in st} == {"n1", "n2"} and len(st) == 6   # 3 x 300m per schedulable node; the 900m DaemonSet pods fit nowhere
    sim = OracleSimulator()
    st, _ = sim.TrySchedulePods(SCENARIOS[0][1], SCENARIOS[0][2], breakOnFailure=True)
    assert st == []   # r1 fails first and breakOnFailure stops the loop


# ---- GPU parity --------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu_engine():
    import __graft_entry__ as g
    g.build()
    from kubernetes_autoscaler_b200.engine import Engine
    e = Engine(device=0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", TRY_CASES, ids=[c[0] for c in TRY_CASES])
def test_gpu_try_schedule_pods_kat(gpu_engine, case):
    _run_try_case(plp.HintingSimulator(gpu_engine), case)


@pytest.mark.gpu
@pytest.mark.parametrize("pod_nodes", HINT_CASES)
def test_gpu_pod_schedules_on_hinted_node_kat(gpu_engine, pod_nodes):
    _run_hint_case(plp.HintingSimulator(gpu_engine), pod_nodes)


@pytest.mark.gpu
@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_gpu_filter_out_schedulable_kat(gpu_engine, case):
    _run_filter_case(lambda f: plp.FilterOutSchedulablePodListProcessor(f, gpu_engine), case)


def _both(gpu_engine, cluster, pods, hints=None, ok=plp.ScheduleAnywhere, brk=False, last_index=0):
    from oracle import pyoracle
    h = plp.Hints()
    for k, v in (hints or {}).items():
        h.Set(("default", k), v)
    x = plp.prepare_try_schedule(cluster, pods, h, ok)
    want = pyoracle.filter_schedulable(x.enc, x.order, x.hint, x.sim_class, x.class_ctrl, x.node_ok, last_index, brk)
    gpu_engine.load(x.enc)
    got = gpu_engine.filter_schedulable(x.order, x.hint, x.sim_class, x.class_ctrl, x.node_ok, last_index, brk)
    assert np.array_equal(got[0], want[0]), (got[0], want[0])
    assert got[1:] == want[1:]
    return want


@pytest.mark.gpu
@pytest.mark.parametrize("scn", SCENARIOS, ids=[s[0] for s in SCENARIOS])
def test_gpu_filter_semantics(gpu_engine, scn):
    _, cluster, pods = scn
    for brk in (False, True):
        for li in (0, len(cluster) - 1, 2 * len(cluster) + 1):   # a lastIndex left by a longer list wraps (plugin_runner.go:81)
            _both(gpu_engine, cluster, pods, brk=brk, last_index=li)
    _both(gpu_engine, cluster, pods, hints={pods[-1].name: cluster[-1].node.name, pods[0].name: cluster[0].node.name})
    _both(gpu_engine, cluster, pods, ok=lambda ni: ni.node.name != cluster[0].node.name)


@pytest.mark.gpu
def test_gpu_filter_synthetic(gpu_engine):
    """Config-3 shaped cluster (2000 -> 300 nodes with free capacity, resident pods, zone / hostname spread) with the
    synthetic pending pods tried on it, in group order and in a scrambled order."""
    from kubernetes_autoscaler_b200 import synth
    from oracle import pyoracle
    enc = synth.generate(3, pods=4_000, templates=8, cluster_nodes=300)
    rng = np.random.default_rng(7)
    for order in (np.arange(enc.P), rng.permutation(enc.P)[:1500]):
        want = pyoracle.filter_schedulable(enc, order)
        gpu_engine.load(enc)
        got = gpu_engine.filter_schedulable(order)
        assert np.array_equal(got[0], want[0])
        assert got[1:] == want[1:]


# ---- scale-down consumer: RemovalSimulator.SimulateNodeRemoval (simulator/cluster_test.go:46-230) ------------------------
def _removal_cases():
    from kubernetes_autoscaler_b200.objects import BuildTestNode as N
    def rs(p):
        p.owner_uid, p.owner_kind = "rs", "ReplicaSet"
        return p
    def topo(name):
        p = rs(BuildTestPod(name, 100, 100000, WithLabels({"app": "topo-app"})))
        p.topology_spread = [TopologySpreadConstraint(1, "kubernetes.io/hostname", LabelSelector({"app": "topo-app"}), min_domains=2)]
        return p
    def tnode(name):
        n = N(name, 1000, 2000000)
        n.labels = {"kubernetes.io/hostname": name}
        return n
    def mk():
        empty = NodeInfo(N("n1", 1000, 2000000))
        drainable = NodeInfo(N("n2", 1000, 2000000), [rs(BuildTestPod("p1", 100, 100000)), rs(BuildTestPod("p2", 100, 100000))])
        nondrain = NodeInfo(N("n3", 1000, 2000000), [BuildTestPod("p3", 100, 100000)])
        full = NodeInfo(N("n4", 1000, 2000000), [BuildTestPod("p4", 1000, 100000)])
        t1 = NodeInfo(tnode("topo-n1"), [topo("p5")])
        t2 = NodeInfo(tnode("topo-n2"), [topo("p6"), BuildTestPod("blocker1", 100, 100000)])
        t3 = NodeInfo(tnode("topo-n3"), [topo("p7"), BuildTestPod("blocker2", 100, 100000)])
        return dict(empty=empty, drainable=drainable, nondrain=nondrain, full=full, t1=t1, t2=t2, t3=t3)
    return mk, [
        ("just an empty node, should be removed", ["empty"], "n1", ("remove", [])),
        ("just a drainable node, but nowhere for pods to go to", ["drainable"], "n2", ("unremovable", "NoPlaceToMovePods")),
        ("drainable node, and a mostly empty node that can take its pods", ["drainable", "nondrain"], "n2", ("remove", ["p1", "p2"])),
        ("drainable node, and a full node that cannot fit anymore pods", ["drainable", "full"], "n2", ("unremovable", "NoPlaceToMovePods")),
        ("4 nodes, 1 empty, 1 drainable", ["empty", "drainable", "full", "nondrain"], "n1", ("remove", [])),
        ("topology spread constraint test - one node should be removable", ["t1", "t2", "t3"], "topo-n1", ("remove", ["p5"])),
        ("candidate not in clusterSnapshot should be marked unremovable", [], "n5", ("unremovable", "NoNodeInfo")),
    ]


def _run_removal_case(make_sim, case):
    from kubernetes_autoscaler_b200.removal import RemovalSimulator
    mk, _ = _removal_cases()
    _, names, node_name, want = case
    nodes = mk()
    cluster = [nodes[n] for n in names]
    r = RemovalSimulator(cluster, False, schedulingSimulator=make_sim())
    to_remove, unremovable = r.SimulateNodeRemoval(node_name, {ni.node.name: True for ni in cluster})
    if want[0] == "remove":
        assert unremovable is None and to_remove.node.name == node_name
        assert [p.name for p in to_remove.pods_to_reschedule] == want[1]
    else:
        assert to_remove is None and unremovable.reason == want[1] and unremovable.node.name == node_name


REMOVAL_CASES = _removal_cases()[1]


@pytest.mark.parametrize("case", REMOVAL_CASES, ids=[c[0] for c in REMOVAL_CASES])
def test_oracle_simulate_node_removal_kat(case):
    _run_removal_case(OracleSimulator, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", REMOVAL_CASES, ids=[c[0] for c in REMOVAL_CASES])
def test_gpu_simulate_node_removal_kat(gpu_engine, case):
    _run_removal_case(lambda: plp.HintingSimulator(gpu_engine), case)


def test_removal_persists_successful_simulations():
    from kubernetes_autoscaler_b200.removal import RemovalSimulator
    mk, _ = _removal_cases()
    nodes = mk()
    cluster = [nodes["drainable"], nodes["nondrain"], nodes["empty"]]
    r = RemovalSimulator(cluster, True, schedulingSimulator=OracleSimulator())
    to_remove, _ = r.SimulateNodeRemoval("n2", {"n1": True, "n3": True})
    assert to_remove is not None and [ni.node.name for ni in cluster] == ["n3", "n1"]
    assert sorted(p.name for ni in cluster for p in ni.pods) == ["p1", "p2", "p3"]
    # hints of the persisted simulation steer the next one (the pods would go back to where they were placed)
    assert r.schedulingSimulator.hints.Get(("default", "p1")) in ("n1", "n3")


# ---- the reference's filter-pass benchmark vector: BenchmarkRunFiltersUntilPassingNode -------------------------------------
# (simulator/clustersnapshot/predicate/plugin_runner_test.go:372-421): 5000 nodes of 10 m CPU with ten 1 m pods each, the
# 5001st node (1000 m) is the only one that takes the 100 m pod; a fresh runner (lastIndex = 0) scans the whole list.
def _benchmark_snapshot():
    cluster = [NodeInfo(_ready_node("n-%d" % i, 10, 1000), [_scheduled("p-%d-%d" % (i, j), 1, 1) for j in range(10)]) for i in range(5000)]
    cluster.append(NodeInfo(_ready_node("n-5000", 1000, 1000)))
    return cluster, [BuildTestPod("p", 100, 1000)]


def _run_benchmark_vector(sim):
    cluster, pods = _benchmark_snapshot()
    for _ in range(2):                      # "lastIndex = 0 // Reset state for each run"
        sim.last_index = 0
        sim.hints = plp.Hints()
        statuses, _ = sim.TrySchedulePods(cluster, pods)
        assert [(s.pod.name, s.node_name) for s in statuses] == [("p", "n-5000")]
        assert sim.last_index == 0          # (5000 + 1) % 5001


def test_oracle_run_filters_until_passing_node_benchmark_vector():
    _run_benchmark_vector(OracleSimulator())


@pytest.mark.gpu
def test_gpu_run_filters_until_passing_node_benchmark_vector(gpu_engine):
    _run_benchmark_vector(plp.HintingSimulator(gpu_engine))
