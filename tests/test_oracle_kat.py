"""Pins the CPU oracle against the reference's own known-answer tests.

Each case transcribes a reference test table entry; file:line given per case.  These run on CPU
(`-m "not gpu"`) and are the gate that lets the oracle arbitrate "bit-exact" for the CUDA path.
"""
import numpy as np
import pytest

from kubernetes_autoscaler_b200.encode import encode
from kubernetes_autoscaler_b200.objects import (BuildTestPod, NodeInfo, WithHostPort, WithLabels,
                                                WithMaxSkew, WithNamespace, makeNode,
                                                makePodEquivalenceGroup)

LABELS = {"app": "estimatee"}


def _pod(cpu, mem, *opts):
    return BuildTestPod("estimatee", cpu, mem, WithNamespace("universe"), WithLabels(LABELS), *opts)


HIGH = makePodEquivalenceGroup(_pod(500, 1000), 10)

# estimator/binpacking_estimator_test.go:90-224 (TestBinpackingEstimate)
CASES = [
    # name, millicores, memory MiB, maxNodes, groups, expect nodes, expect pods, expected sched per group
    ("simple resource-based binpacking (:91)", 350 * 3 - 50, 2 * 1000, 0,
     [makePodEquivalenceGroup(_pod(350, 1000), 10)], 5, 10, None),
    ("pods-per-node bound binpacking (:107)", 10000, 20000, 0,
     [makePodEquivalenceGroup(_pod(10, 100), 20)], 2, 20, None),
    ("hostport conflict forces pod-per-node (:123)", 1000, 5000, 0,
     [makePodEquivalenceGroup(_pod(200, 1000, WithHostPort(5555)), 8)], 8, 8, None),
    ("limiter cuts binpacking (:140)", 1000, 5000, 5,
     [makePodEquivalenceGroup(_pod(500, 1000), 20)], 5, 10, None),
    ("decreasing ordered pods are processed first (:157)", 1000, 5000, 5,
     [makePodEquivalenceGroup(_pod(50, 1000), 10), HIGH], 5, 10, [0, 10]),
    ("hostname topology spreading with maxSkew=2 forces 2 pods/node (:175)", 1000, 5000, 0,
     [makePodEquivalenceGroup(_pod(200, 200, WithMaxSkew(2, "kubernetes.io/hostname", 1)), 8)], 4, 8, None),
    ("zonal topology spreading with maxSkew=2 only allows 2 pods to schedule (:192)", 1000, 5000, 0,
     [makePodEquivalenceGroup(_pod(20, 100, WithMaxSkew(2, "topology.kubernetes.io/zone", 1)), 8)], 1, 2, None),
    ("hostname topology spreading maxSkew=1 minDomains=3 schedules retroactively (:209)", 1000, 5000, 0,
     [makePodEquivalenceGroup(_pod(20, 100, WithMaxSkew(1, "kubernetes.io/hostname", 3)), 12)], 3, 12, None),
]


def _fixture(millicores, memory, pods_per_node, groups):
    """binpacking_estimator_test.go:228-239: snapshot holds `oldnode`; template in zone-mars."""
    cluster = [NodeInfo(makeNode(100, 100, 10, "oldnode", "zone-jupiter"))]
    template = [NodeInfo(makeNode(millicores, memory, pods_per_node, "template", "zone-mars"))]
    return encode(cluster, template, groups)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_binpacking_estimate_kat(oracle, case):
    name, cpu, mem, max_nodes, groups, exp_nodes, exp_pods, exp_sched = case
    enc = _fixture(cpu, mem, 10, groups)
    nodes, pods, sched, order, _ = oracle.estimate(enc, 0, max_nodes=max_nodes)
    assert (nodes, pods) == (exp_nodes, exp_pods)
    if exp_sched is not None:
        # expectProcessedPods == highResourcePodGroup.Pods (:172): only group 1 scheduled, and first
        assert list(sched) == exp_sched
        assert order[0] == 1


def test_kat8_oldnode_counted(oracle):
    """SURVEY §8c trace of KAT 8: pod 2 falls back to `oldnode` (list index 0), which is then counted."""
    case = CASES[7]
    enc = _fixture(case[1], case[2], 10, case[4])
    nodes, pods, sched, order, placements = oracle.estimate(enc, 0, max_nodes=0)
    assert nodes == 3 and pods == 12
    assert 0 in set(placements.tolist())  # oldnode received pods via binpacking_estimator.go:190-205
    assert placements[0] == 1 and placements[1] == 0


def test_benchmark_vector(oracle):
    """BenchmarkBinpackingEstimate (:249-296): 50 000 + 1000 pods -> 2595 nodes / 51 000 pods."""
    groups = [makePodEquivalenceGroup(_pod(50, 100), 50000), makePodEquivalenceGroup(_pod(95, 190), 1000)]
    enc = _fixture(1000, 5000, 100, groups)
    nodes, pods, sched, order, _ = oracle.estimate(enc, 0, max_nodes=3000)
    assert (nodes, pods) == (2595, 51000)
    assert order == [1, 0]  # the 95 m group scores higher


def test_decreasing_pod_orderer(oracle):
    """estimator/decreasing_pod_orderer_test.go:28-65: node 4000m / 5000 MiB... order p4,p3,p2,p1."""
    from kubernetes_autoscaler_b200.objects import BuildTestNode
    node = BuildTestNode("node-1", 1000, 1000)  # :29 BuildTestNode("node1", 1000, 1000) capacity only
    pods = [BuildTestPod("p1", 1, 1), BuildTestPod("p2", 2, 2), BuildTestPod("p3", 3, 3), BuildTestPod("p4", 4, 4)]
    groups = [makePodEquivalenceGroup(p, 5) for p in pods]
    enc = encode([], [NodeInfo(node)], groups)
    scores = [oracle.pod_score(enc, int(enc.arrays["pend_spec"][enc.arrays["group_off"][g]]), 0) for g in range(4)]
    assert scores == sorted(scores) and len(set(scores)) == 4
    _, _, _, order, _ = oracle.estimate(enc, 0)
    assert order == [3, 2, 1, 0]


def test_get_min_limit_table(oracle):
    """estimator/threshold_based_limiter_test.go:180-203 TestMinLimit (int table, verbatim)."""
    table = [(-10, 10, -1), (-10, 0, -1), (-10, -10, -1), (0, 0, 0), (0, 10, 10), (5, 10, 5)]
    for base, target, want in table:
        assert oracle.get_min_limit(base, target) == want, (base, target)


def test_threshold_based_limiter(oracle):
    """estimator/threshold_based_limiter_test.go:56-160, the node-count cases (duration cases are
    wall-clock and stay in the Go shim)."""
    # "no limiting happens" (:66): no thresholds -> 3 allows
    assert oracle.limiter_grants([], 3) == 3
    # "sequence of additions works until the threshold is hit" (:87): static 3 -> allow x3, deny
    assert oracle.limiter_grants([3], 4) == 3
    # "binpacking is stopped if at least one threshold has negative max nodes limit" (:98)
    assert oracle.limiter_grants([-1, 10], 1) == 0
    # "node counter is reset" (:120): static 2 -> allow, allow, deny; after reset allow again
    assert oracle.limiter_grants([2], 3) == 2
    assert oracle.limiter_grants([2], 1) == 1


def test_capacity_thresholds(oracle):
    """cluster_capacity_threshold_test.go:33-56 and sng_capacity_threshold_test.go:38-77 (verbatim)."""
    assert oracle.cluster_capacity_limit(True, 10, 5) == 5      # returns available capacity
    assert oracle.cluster_capacity_limit(True, 0, 10) == 0      # unlimited
    assert oracle.cluster_capacity_limit(True, 5, 10) == -1     # no capacity
    assert oracle.cluster_capacity_limit(True, -5, 0) == -1     # negative limit
    assert oracle.cluster_capacity_limit(False, 10, 5) == 0     # nil context (threshold.go contract)
    # (current group first, then the similar node groups)
    assert oracle.sng_capacity_limit(True, [20, 10, 100, 5], [10, 5, 50, 3]) == 67
    assert oracle.sng_capacity_limit(True, [5, 10, 10, 0], [10, 5, 11, 5]) == 5
    assert oracle.sng_capacity_limit(True, [5, 10, 100], [5, 10, 100]) == -1
    assert oracle.sng_capacity_limit(True, [5, 10, 100, 0], [10, 11, 111, 5]) == -1
    assert oracle.sng_capacity_limit(False, [10], [5]) == 0


# simulator/clustersnapshot/predicate/plugin_runner_test.go:43-196 (TestRunFiltersOnNode, default scheduler config):
# (scheduled pods on n1000, test pod, failing plugin or None)
def _runner_cases():
    from kubernetes_autoscaler_b200.objects import BuildTestNode, WithNodeNamesAffinity
    p450, p600, p8000, p500 = (BuildTestPod("p450", 450, 500000), BuildTestPod("p600", 600, 500000),
                               BuildTestPod("p8000", 8000, 0), BuildTestPod("p500", 500, 500000))
    aff = BuildTestPod("pod_with_affinity", 500, 500, WithNodeNamesAffinity("n1000"))
    bad_aff = BuildTestPod("pod_with_affinity", 500, 500, WithNodeNamesAffinity("non-existing-node"))
    n1000 = lambda pods: NodeInfo(BuildTestNode("n1000", 1000, 2000000), pods)
    FIT, PREFILTER = 7, 1   # CAE_R_FIT ("NodeResourcesFit": Insufficient cpu), CAE_R_PREFILTER_NODEAFFINITY ("PreFilter filtered the Node out")
    return [
        ("default - other pod - insuficient cpu (:85)", n1000([p450]), p600, FIT),
        ("default - other pod - ok (:95)", n1000([p450]), p500, 0),
        ("default - empty - insuficient cpu (:102)", n1000([]), p8000, FIT),
        ("default - empty - ok (:112)", n1000([]), p600, 0),
        ("default - affinity on existing node - ok (:121)", n1000([]), aff, 0),
        ("default - affinity on non-existing node - error (:128)", n1000([]), bad_aff, PREFILTER),
    ]


RUNNER_CASES = _runner_cases()


@pytest.mark.parametrize("case", RUNNER_CASES, ids=[c[0] for c in RUNNER_CASES])
def test_run_filters_on_node_kat(oracle, case):
    """The node under test is the 'template' (CheckPredicates on a node that joined the snapshot); its scheduled pods are the
    NodeInfo's pods."""
    _, node, pod, want = case
    enc = encode([], [node], [makePodEquivalenceGroup(pod, 1)])
    reasons, _ = oracle.feasibility_dense(enc)
    assert int(reasons[0][0]) == want


# plugin_runner_test.go:198-294 (TestRunFilterUntilPassingNode, default config): nodes n1000, n2000
@pytest.mark.parametrize("cpu,expected", [(900, {"n1000", "n2000"}), (1900, {"n2000"}), (2100, set())],
                         ids=["default - small pod - no error (:231)", "default - medium pod - no error (:237)",
                              "default - large pod - insufficient cpu (:243)"])
def test_run_filters_until_passing_node_kat(oracle, cpu, expected):
    from kubernetes_autoscaler_b200.objects import BuildTestNode
    cluster = [NodeInfo(BuildTestNode("n1000", 1000, 2000000)), NodeInfo(BuildTestNode("n2000", 2000, 2000000))]
    enc = encode(cluster, [], [makePodEquivalenceGroup(BuildTestPod("p", cpu, 1000), 1)])
    assigned, _, _ = oracle.filter_schedulable(enc, [0])
    if expected:
        assert cluster[int(assigned[0])].node.name in expected
    else:
        assert int(assigned[0]) == -1


def test_debug_info_kat(oracle):
    """plugin_runner_test.go:296-324 (TestDebugInfo, default config): a node with a NoSchedule and a NoExecute taint rejects a
    pod without tolerations; the failing predicate is TaintToleration with its reference reason string."""
    from kubernetes_autoscaler_b200 import capi
    from kubernetes_autoscaler_b200.objects import BuildTestNode, Taint
    n1 = BuildTestNode("n1", 1000, 2000000)
    n1.taints = [Taint("SomeTaint", "WhyNot?", "NoSchedule"), Taint("RandomTaint", "JustBecause", "NoExecute")]
    enc = encode([], [NodeInfo(n1)], [makePodEquivalenceGroup(BuildTestPod("p1", 0, 0), 1)])
    reasons, _ = oracle.feasibility_dense(enc)
    plugin, reason = capi.REASON_PLUGIN[int(reasons[0][0])]
    assert (plugin, reason) == ("TaintToleration", "node(s) had untolerated taint(s)")
