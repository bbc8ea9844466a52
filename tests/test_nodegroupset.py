"""processors/nodegroupset/balancing_processor_test.go:111-262 and orchestrator.go:757-812 on the host mirror."""
import pytest

from kubernetes_autoscaler_b200.estimator import NodeGroupInfo
from kubernetes_autoscaler_b200.nodegroupset import (BalanceScaleUpBetweenGroups, ComputeSimilarNodeGroups,
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
