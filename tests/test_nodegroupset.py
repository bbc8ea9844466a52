"""processors/nodegroupset/balancing_processor_test.go:111-262 and orchestrator.go:757-812 on the host mirror."""
import pytest

from kubernetes_autoscaler_b200.estimator import NodeGroupInfo
from nodegroupset_harness import (BalanceScaleUpBetweenGroups, ComputeSimilarNodeGroups,
                                                     matchingSchedulablePodGroups)


def _ng(name, mx, size):      # provider.AddNodeGroup(id, min, max, size)
    return NodeGroupInfo(name, max_size=mx, target_size=size)


def test_balance_single_group():
    g = [_ng("ng1", 10, 1)]
    r = BalanceScaleUpBetweenGroups(g, 1)
    assert len(r) == 1 and r[0].new_size == 2
    r = BalanceScaleUpBetweenGroups(g, 4)
    assert len(r) == 1 and r[0].new_size == 5


def test_balance_under_max_size():
    g = [_ng("ng1", 10, 1), _ng("ng2", 10, 3), _ng("ng3", 10, 5), _ng("ng4", 10, 5)]
    r = BalanceScaleUpBetweenGroups(g, 1)
    assert [(i.group.id, i.new_size) for i in r] == [("ng1", 2)]
    r = BalanceScaleUpBetweenGroups(g, 2)
    assert [(i.group.id, i.new_size) for i in r] == [("ng1", 3)]
    r = BalanceScaleUpBetweenGroups(g, 4)                       # divisible
    assert sorted((i.group.id, i.new_size) for i in r) == [("ng1", 4), ("ng2", 4)]
    r = BalanceScaleUpBetweenGroups(g, 5)                       # non-divisible: 4 and 5
    assert sorted(i.group.id for i in r) == ["ng1", "ng2"] and sum(i.new_size for i in r) == 9
    assert all(i.new_size in (4, 5) for i in r)
    r = BalanceScaleUpBetweenGroups(g, 10)                      # all groups, divisible
    assert len(r) == 4 and all(i.new_size == 6 for i in r)


def test_balance_hitting_max_size():
    ngs = {"ng1": _ng("ng1", 1, 1), "ng2": _ng("ng2", 3, 1), "ng3": _ng("ng3", 10, 3), "ng4": _ng("ng4", 7, 5), "ng5": _ng("ng5", 3, 6)}
    get = lambda *names: [ngs[n] for n in names]
    as_map = lambda r: {i.group.id: i.new_size for i in r}
    assert BalanceScaleUpBetweenGroups(get("ng1"), 1) == []                      # just one maxed out group
    assert as_map(BalanceScaleUpBetweenGroups(get("ng1", "ng2"), 1)) == {"ng2": 2}
    assert as_map(BalanceScaleUpBetweenGroups(get("ng1", "ng2"), 5)) == {"ng2": 3}   # capped to capacity
    assert as_map(BalanceScaleUpBetweenGroups(get("ng2", "ng3"), 4)) == {"ng2": 3, "ng3": 5}
    assert as_map(BalanceScaleUpBetweenGroups(get("ng2", "ng3", "ng4"), 9)) == {"ng2": 3, "ng3": 8, "ng4": 7}
    assert as_map(BalanceScaleUpBetweenGroups(get("ng2", "ng3", "ng4"), 900)) == {"ng2": 3, "ng3": 10, "ng4": 7}
    assert as_map(BalanceScaleUpBetweenGroups(get("ng2", "ng5"), 1)) == {"ng2": 2}   # one group exceeds its max
    with pytest.raises(ValueError):
        BalanceScaleUpBetweenGroups([], 1)


def test_compute_similar_node_groups():
    sched = {"a": [0, 2], "b": [0, 1, 2], "c": [2], "d": []}
    assert matchingSchedulablePodGroups(sched["a"], sched["b"]) and not matchingSchedulablePodGroups(sched["b"], sched["a"])
    assert ComputeSimilarNodeGroups("a", ["b", "c", "missing"], sched) == ["b"]
    assert ComputeSimilarNodeGroups("c", ["a", "b"], sched) == ["a", "b"]
    assert ComputeSimilarNodeGroups("d", ["a"], sched) == []                      # nothing schedulable on the main group
    assert ComputeSimilarNodeGroups("a", ["b"], sched, balance_similar_node_groups=False) == []
    assert ComputeSimilarNodeGroups("a", ["b"], sched, zero_or_max_node_scaling=True) == []


# ---- processors/nodegroupset/compare_nodegroups_test.go:41-197 ----------------------------------------------------------------
from nodegroupset_harness import CreateGenericNodeInfoComparator, FindSimilarNodeGroups  # noqa: E402
from kubernetes_autoscaler_b200.objects import BuildTestNode, BuildTestPod, NodeInfo  # noqa: E402
from kubernetes_autoscaler_b200.snapshotz import quantity_value  # noqa: E402


def _similar(n1, n2, want, pods1=(), pods2=(), extra=()):
    cmp = CreateGenericNodeInfoComparator(extra)
    assert cmp(NodeInfo(n1, list(pods1)), NodeInfo(n2, list(pods2))) is want


def test_identical_nodes_similar():
    _similar(BuildTestNode("node1", 1000, 2000), BuildTestNode("node2", 1000, 2000), True)


def test_nodes_similar_various_requirements():
    n1 = BuildTestNode("node1", 1000, 2000)
    n2 = BuildTestNode("node2", 1000, 2000)
    n2.capacity["cpu"] = 1001                                   # different CPU capacity
    _similar(n1, n2, False)
    n3 = BuildTestNode("node3", 1000, 2000)
    n3.allocatable["cpu"] = 999                                 # slightly different allocatable
    _similar(n1, n3, True)
    n4 = BuildTestNode("node4", 1000, 2000)
    n4.allocatable["cpu"] = 500                                 # significantly different allocatable
    _similar(n1, n4, False)
    n5 = BuildTestNode("node5", 1000, 2000)
    n5.capacity["nvidia.com/gpu"] = n5.allocatable["nvidia.com/gpu"] = 1   # one with GPU, one without
    _similar(n1, n5, False)


def test_nodes_similar_various_requirements_and_pods():
    n1, p1 = BuildTestNode("node1", 1000, 2000), BuildTestPod("pod1", 500, 1000)
    n2 = BuildTestNode("node2", 1000, 2000)
    n2.allocatable["cpu"], n2.allocatable["memory"] = 500, 1000  # different allocatable, but same free
    _similar(n1, n2, False, [p1], [])
    _similar(n1, BuildTestNode("node3", 1000, 2000), True, [p1], [BuildTestPod("pod3", 500, 1000)])
    n4 = BuildTestNode("node4", 1000, 2000)
    n4.allocatable["cpu"] = 999                                 # similar allocatable, similar pods
    _similar(n1, n4, True, [p1], [BuildTestPod("pod4", 501, 1001)])


def test_nodes_similar_various_memory_requirements():
    n1 = BuildTestNode("node1", 1000, 1000)
    n2 = BuildTestNode("node2", 1000, 1000)
    n2.capacity["memory"] = int(1000 - (1000 * 0.015) + 1)
    _similar(n1, n2, True)
    n3 = BuildTestNode("node3", 1000, 1000)
    n3.capacity["memory"] = int(1000 - (1000 * 0.015) - 1)
    _similar(n1, n3, False)


@pytest.mark.parametrize("q1,q2,q3", [("16116152Ki", "15944120Ki", "16438475Ki"), ("259970052Ki", "257217528Ki", "265169453Ki")],
                         ids=["m5.xlarge", "m5.16xlarge"])
def test_nodes_similar_large_memory(q1, q2, q3):
    n1 = BuildTestNode("node1", 1000, quantity_value(q1))
    _similar(n1, BuildTestNode("node2", 1000, quantity_value(q2)), True)    # another zone's instance of the same type
    _similar(n1, BuildTestNode("node3", 1000, quantity_value(q3)), False)   # q1 * 1.02


def test_nodes_similar_various_labels():
    extra = ["example.com/ready"]
    n1, n2 = BuildTestNode("node1", 1000, 2000), BuildTestNode("node2", 1000, 2000)
    n1.labels.update({"test-label": "test-value", "character": "winnie the pooh"})
    n2.labels["test-label"] = "test-value"
    _similar(n1, n2, False, extra=extra)                        # missing character label
    n2.labels["character"] = "winnie the pooh"
    _similar(n1, n2, True, extra=extra)
    n1.labels["kubernetes.io/hostname"], n2.labels["kubernetes.io/hostname"] = "node1", "node2"
    _similar(n1, n2, True, extra=extra)
    n1.labels["failure-domain.beta.kubernetes.io/zone"], n2.labels["failure-domain.beta.kubernetes.io/zone"] = "mars-olympus-mons1-b", "us-houston1-a"
    _similar(n1, n2, True, extra=extra)
    n1.labels["beta.kubernetes.io/fluentd-ds-ready"], n2.labels["beta.kubernetes.io/fluentd-ds-ready"] = "true", "false"
    _similar(n1, n2, True, extra=extra)
    del n2.labels["beta.kubernetes.io/fluentd-ds-ready"]
    _similar(n1, n2, True, extra=extra)
    n1.labels["example.com/ready"], n2.labels["example.com/ready"] = "true", "false"
    _similar(n1, n2, True, extra=extra)


def test_find_similar_node_groups():
    infos = {"a": NodeInfo(BuildTestNode("ta", 1000, 2000)), "b": NodeInfo(BuildTestNode("tb", 1000, 2000)),
             "c": NodeInfo(BuildTestNode("tc", 2000, 2000))}
    assert FindSimilarNodeGroups("a", infos) == ["b"]
    assert FindSimilarNodeGroups("c", infos) == []
