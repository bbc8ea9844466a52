"""Decision-level known-answer tests transcribed from the reference's own orchestrator tests
(cluster-autoscaler/core/scaleup/orchestrator/orchestrator_test.go): TestScaleUpOK (:75), TestMixedScaleUp (:101), the
three GPU-pool cases (:853-954), TestBinpackingLimiter (:1349) and TestScaleUpBalanceGroups (:1622).

The reference drives ScaleUp with a reporting mock expander that records the options it was offered and picks a given
one; here the same shape: `expander_strategy` records + picks.  Each case runs on the CPU oracle double (always) and, under
-m gpu, through the real engine — the two must give the reference's expected option list, final option and pod lists."""
import pytest

from kubernetes_autoscaler_b200.estimator import NodeGroupInfo
from kubernetes_autoscaler_b200.objects import BuildTestNode, BuildTestPod, NodeInfo
from scaleup_harness import AutoscalingOptions, ScaleUpOrchestrator, ScaleUpSuccessful
from test_scaleup_orchestrator import OracleEngine

MiB = 1 << 20
GPU = "nvidia.com/gpu"


def _node(name, cpu, mem, gpu=0):
    n = BuildTestNode(name, cpu, mem)
    if gpu:
        n.allocatable[GPU] = gpu
        n.capacity[GPU] = gpu
    return n


def _pod(name, cpu, mem, gpu=0):
    p = BuildTestPod(name, cpu, mem)
    if gpu:
        p.requests[GPU] = gpu
    return p


class ReportingStrategy:
    """MockReportingStrategy (core/test/common.go): remembers the options it was offered, returns the requested one."""

    def __init__(self, choose):
        self.choose = choose
        self.offered = None

    def __call__(self, options):
        self.offered = [(o.node_group, o.node_count) for o in options]
        for o in options:
            if (o.node_group, o.node_count) == self.choose:
                return o
        return None


def _run(engine, nodes, resident, extra, choose, options=None, groups_meta=None, expected_options=None):
    """nodes: [(name, cpu, mem, gpu, group)], resident: [(pod, node)], extra: pending pods.  Node groups: min 1, max 10,
    target = number of nodes of the group (simpleScaleUpTest's provider set-up); template of a group = its first node."""
    by_group = {}
    for name, cpu, mem, gpu, grp in nodes:
        by_group.setdefault(grp, []).append(_node(name, cpu, mem, gpu))
    pods_on = {}
    for p, n in resident:
        pods_on.setdefault(n, []).append(p)
    cluster = [NodeInfo(n, pods_on.get(n.name, [])) for ns in by_group.values() for n in ns]
    node_infos = {g: NodeInfo(_node(ns[0].name + "-template", ns[0].allocatable["cpu"], ns[0].allocatable["memory"],
                                    ns[0].allocatable.get(GPU, 0))) for g, ns in by_group.items()}
    ngs = [NodeGroupInfo(g, (groups_meta or {}).get(g, (10, len(ns)))[0], (groups_meta or {}).get(g, (10, len(ns)))[1]) for g, ns in by_group.items()]
    strat = ReportingStrategy(choose) if choose else None
    orch = ScaleUpOrchestrator(options or AutoscalingOptions(), engine=engine)
    status = orch.ScaleUp(extra, cluster, node_infos, ngs, expander_strategy=strat)
    if expected_options is not None:
        assert sorted(strat.offered) == sorted(expected_options)
    return status, orch


def _names(ps):
    return sorted(p.name for p in ps)


def _cases(engine):
    # TestScaleUpOK (:75-99)
    st, _ = _run(engine, [("n1", 100, 100, 0, "ng1"), ("n2", 1000, 1000, 0, "ng2")],
                 [(_pod("p1", 80, 0), "n1"), (_pod("p2", 800, 0), "n2")], [_pod("p-new", 500, 0)], ("ng2", 1))
    assert st.result == ScaleUpSuccessful
    assert [(i.group.id, i.new_size - i.current_size) for i in st.scale_up_infos] == [("ng2", 1)]
    assert _names(st.pods_triggered_scale_up) == ["p-new"]
    # TestMixedScaleUp (:101-131): triggering, remaining and awaiting pods
    st, _ = _run(engine, [("n1", 100, 1000, 0, "ng1"), ("n2", 1000, 100, 0, "ng2")],
                 [(_pod("p1", 80, 0), "n1"), (_pod("p2", 800, 0), "n2")],
                 [_pod("triggering", 900, 0), _pod("remaining", 2000, 0), _pod("awaiting", 0, 200)], ("ng2", 1))
    assert [(i.group.id, i.new_size - i.current_size) for i in st.scale_up_infos] == [("ng2", 1)]
    assert _names(st.pods_triggered_scale_up) == ["triggering"]
    assert _names(st.pods_remain_unschedulable) == ["remaining"]
    assert _names(st.pods_await_evaluation) == ["awaiting"]
    # TestWillConsiderGpuAndStandardPoolForPodWhichDoesNotRequireGpu (:853-883)
    opts = AutoscalingOptions(max_nodes_total=100)
    two = [("gpu-node-1", 2000, 1000 * MiB, 1, "gpu-pool"), ("std-node-1", 2000, 1000 * MiB, 0, "std-pool")]
    res = [(_pod("gpu-pod-1", 2000, 1000 * MiB, 1), "gpu-node-1"), (_pod("std-pod-1", 2000, 1000 * MiB), "std-node-1")]
    st, _ = _run(engine, two, res, [_pod("extra-std-pod", 2000, 1000 * MiB)], ("std-pool", 1), opts,
                 expected_options=[("std-pool", 1), ("gpu-pool", 1)])
    assert [(i.group.id, i.new_size - i.current_size) for i in st.scale_up_infos] == [("std-pool", 1)]
    assert _names(st.pods_triggered_scale_up) == ["extra-std-pod"]
    # TestWillConsiderOnlyGpuPoolForPodWhichDoesRequiresGpu (:885-914)
    st, _ = _run(engine, two, res, [_pod("extra-gpu-pod", 2000, 1000 * MiB, 1)], ("gpu-pool", 1), opts, expected_options=[("gpu-pool", 1)])
    assert [(i.group.id, i.new_size - i.current_size) for i in st.scale_up_infos] == [("gpu-pool", 1)]
    assert _names(st.pods_triggered_scale_up) == ["extra-gpu-pod"]
    # TestWillConsiderAllPoolsWhichFitTwoPodsRequiringGpus (:916-954)
    four = [("gpu-1-node-1", 2000, 1000 * MiB, 1, "gpu-1-pool"), ("gpu-2-node-1", 2000, 1000 * MiB, 2, "gpu-2-pool"),
            ("gpu-4-node-1", 2000, 1000 * MiB, 4, "gpu-4-pool"), ("std-node-1", 2000, 1000 * MiB, 0, "std-pool")]
    res4 = [(_pod("gpu-pod-1", 2000, 1000 * MiB, 1), "gpu-1-node-1"), (_pod("gpu-pod-2", 2000, 1000 * MiB, 2), "gpu-2-node-1"),
            (_pod("gpu-pod-3", 2000, 1000 * MiB, 4), "gpu-4-node-1"), (_pod("std-pod-1", 2000, 1000 * MiB), "std-node-1")]
    extra = [_pod("extra-gpu-pod-%d" % i, 1, 1 * MiB, 1) for i in (1, 2, 3)]
    st, _ = _run(engine, four, res4, extra, ("gpu-1-pool", 3), opts,
                 expected_options=[("gpu-1-pool", 3), ("gpu-2-pool", 2), ("gpu-4-pool", 1)])
    assert [(i.group.id, i.new_size - i.current_size) for i in st.scale_up_infos] == [("gpu-1-pool", 3)]
    assert _names(st.pods_triggered_scale_up) == ["extra-gpu-pod-1", "extra-gpu-pod-2", "extra-gpu-pod-3"]


def _binpacking_limiter_case(engine):
    """TestBinpackingLimiter (:1349-1409): MockBinpackingLimiter stops after the first option; without it there are two."""
    class StopAfterFirst:
        def StopBinpacking(self, options):
            return len(options) == 1
    nodes = [("n1", 1000, 1000, 0, "ng1"), ("n2", 100000, 100000, 0, "ng2")]
    for limiter, n_opts in ((StopAfterFirst(), 1), (None, 2)):
        by = {"ng1": (10, 1), "ng2": (10, 1)}
        cluster = [NodeInfo(_node(n, c, m)) for n, c, m, _, _ in nodes]
        node_infos = {g: NodeInfo(_node(n + "-t", c, m)) for n, c, m, _, g in nodes}
        ngs = [NodeGroupInfo(g, *by[g]) for g in ("ng1", "ng2")]
        seen = ReportingStrategy(None)
        seen.choose = None

        def pick(options, seen=seen):
            seen.offered = [(o.node_group, o.node_count) for o in options]
            return options[0]
        orch = ScaleUpOrchestrator(AutoscalingOptions(), engine=engine)
        st = orch.ScaleUp([_pod("p-new", 500, 0)], cluster, node_infos, ngs, expander_strategy=pick, binpacking_limiter=limiter)
        assert st.result == ScaleUpSuccessful
        assert len(seen.offered) == n_opts


def _balance_groups_case(engine):
    """TestScaleUpBalanceGroups (:1622-1729): two pods, four similar groups; ng2 and ng3 end at 2 nodes each."""
    cfg = {"ng1": (1, 1), "ng2": (2, 1), "ng3": (5, 1), "ng4": (5, 3)}     # max, size
    cluster, node_infos, ngs = [], {}, []
    for gid, (mx, size) in cfg.items():
        for i in range(size):
            name = "%s-node-%d" % (gid, i)
            p = _pod("%s-pod-%d" % (gid, i), 80, 0)
            cluster.append(NodeInfo(_node(name, 100, 1000), [p]))
        node_infos[gid] = NodeInfo(_node(gid + "-template", 100, 1000))
        ngs.append(NodeGroupInfo(gid, mx, size))
    orch = ScaleUpOrchestrator(AutoscalingOptions(balance_similar_node_groups=True), engine=engine)
    pods = [_pod("test-pod-%d" % i, 80, 0) for i in range(2)]
    st = orch.ScaleUp(pods, cluster, node_infos, ngs, expander_strategy=lambda options: next(o for o in options if o.node_group == "ng2"))
    assert st.result == ScaleUpSuccessful
    sizes = {i.group.id: i.new_size for i in st.scale_up_infos}
    assert sizes.get("ng2") == 2 and sizes.get("ng3") == 2


def test_reference_decision_kats_on_the_oracle_double():
    _cases(OracleEngine())
    _binpacking_limiter_case(OracleEngine())
    _balance_groups_case(OracleEngine())


def test_balance_regression_group_fills_mid_sweep():
    """A(max 10), B(max 1), C(max 10) at size 0, five new nodes -> A=2, B=1, C=2 (Go reads the slot after
    the swap, balancing_processor.go:150-170)."""
    from nodegroupset_harness import BalanceScaleUpBetweenGroups
    infos = BalanceScaleUpBetweenGroups([NodeGroupInfo("A", 10, 0), NodeGroupInfo("B", 1, 0), NodeGroupInfo("C", 10, 0)], 5)
    assert {i.group.id: i.new_size for i in infos} == {"A": 2, "B": 1, "C": 2}


@pytest.mark.gpu
def test_reference_decision_kats_on_the_engine():
    import __graft_entry__ as g
    g.build()
    from kubernetes_autoscaler_b200.engine import Engine
    eng = Engine(device=0)
    try:
        _cases(eng)
        _binpacking_limiter_case(eng)
        _balance_groups_case(eng)
    finally:
        eng.close()
