"""Expander filters pinned on the reference's own tests (expander/waste/waste_test.go:83-133, mostpods/mostpods_test.go,
leastnodes/leastnodes_test.go:27-110): the oracle's restatement AND the product's host chain (cae_expander_chain needs the
library, not a GPU) must both give the expected option sets."""
import numpy as np
import pytest

from kubernetes_autoscaler_b200.encode import encode
from kubernetes_autoscaler_b200.engine import expander_chain
from kubernetes_autoscaler_b200.objects import BuildTestPod, Node, NodeInfo, makePodEquivalenceGroup

LEAST_WASTE, MOST_PODS, LEAST_NODES = 0, 1, 2
CPU_PER_POD, MEM_PER_POD = 500, 1000 * 1024 * 1024


def _node_info(name, cpu, mem, pods=100):
    cap = {"cpu": cpu, "memory": mem, "pods": pods}
    return NodeInfo(Node(name=name, capacity=dict(cap), allocatable=dict(cap)))


def _both(oracle, enc, chain, nc, pc, sched):
    nc, pc, sched = np.asarray(nc, np.int32), np.asarray(pc, np.int32), np.asarray(sched, np.int32).reshape(enc.T, max(enc.E, 1))
    omask, owaste = oracle.expander(enc, chain, nc, pc, sched)
    hmask = expander_chain(chain, nc, pc, owaste)
    assert np.array_equal(omask, hmask)
    return [t for t in range(enc.T) if omask[t]]


def test_least_waste_reference_sequence(oracle):
    pod = BuildTestPod("p", CPU_PER_POD, MEM_PER_POD)
    groups = [makePodEquivalenceGroup(pod, 1)]
    balanced = _node_info("balanced", 16 * CPU_PER_POD, 16 * MEM_PER_POD)
    highmem = _node_info("highmem", 16 * CPU_PER_POD, 32 * MEM_PER_POD)
    lowcpu = _node_info("lowcpu", 8 * CPU_PER_POD, 16 * MEM_PER_POD)
    # without any pods, one node info
    enc = encode([], [balanced], groups)
    assert _both(oracle, enc, [LEAST_WASTE], [1], [0], [0]) == [0]
    # one pod, one node info
    assert _both(oracle, enc, [LEAST_WASTE], [1], [1], [1]) == [0]
    # one pod, two node infos, one has lots of RAM
    enc = encode([], [balanced, highmem], groups)
    assert _both(oracle, enc, [LEAST_WASTE], [1, 1], [1, 1], [1, 1]) == [0]
    # three node infos, one with less CPU wins
    enc = encode([], [balanced, highmem, lowcpu], groups)
    assert _both(oracle, enc, [LEAST_WASTE], [1, 1, 1], [1, 1, 1], [1, 1, 1]) == [2]


def test_most_pods_reference_sequence(oracle):
    groups = [makePodEquivalenceGroup(BuildTestPod("p", 100, 100), 1)]
    tmpl = [_node_info("t%d" % i, 1000, 1 << 30) for i in range(3)]
    enc = encode([], tmpl[:1], groups)
    assert _both(oracle, enc, [MOST_PODS], [1], [0], [0]) == [0]                     # EO0 alone
    enc = encode([], tmpl[:2], groups)
    assert _both(oracle, enc, [MOST_PODS], [1, 1], [0, 1], [0, 1]) == [1]            # EO1 has a pod
    enc = encode([], tmpl, groups)
    assert _both(oracle, enc, [MOST_PODS], [1, 1, 1], [0, 1, 1], [0, 1, 1]) == [1, 2]  # EO1, EO1b tie


@pytest.mark.parametrize("counts,want", [
    ([], []), ([0], []), ([2], [0]), ([2, 1], [1]), ([6, 2, 2], [1, 2]), ([8, 8, 8], [0, 1, 2]),
], ids=["no options", "no valid options", "1 valid option", "2 valid options, not equal", "3 valid options, 2 equal",
        "3 valid options, all equal"])
def test_least_nodes_reference_table(oracle, counts, want):
    groups = [makePodEquivalenceGroup(BuildTestPod("p", 100, 100), 1)]
    tmpl = [_node_info("t%d" % i, 1000, 1 << 30) for i in range(max(len(counts), 1))]
    enc = encode([], tmpl, groups)
    nc = counts if counts else [0]
    assert _both(oracle, enc, [LEAST_NODES], nc, [1] * len(nc), [1] * len(nc)) == want


def test_chain_stops_at_a_single_survivor(oracle):
    """factory/chain.go:36-45: filters run in order until one option is left; ties fall through to the next filter."""
    pod = BuildTestPod("p", CPU_PER_POD, MEM_PER_POD)
    groups = [makePodEquivalenceGroup(pod, 4)]
    a, b, c = (_node_info(n, 16 * CPU_PER_POD, 16 * MEM_PER_POD) for n in "abc")
    enc = encode([], [a, b, c], groups)
    # a and b waste the same (2 pods on 1 node), c wastes more (1 pod); most-pods cannot split a/b; least-nodes does
    assert _both(oracle, enc, [LEAST_WASTE], [1, 1, 1], [2, 2, 1], [2, 2, 1]) == [0, 1]
    assert _both(oracle, enc, [LEAST_WASTE, MOST_PODS], [1, 1, 1], [2, 2, 1], [2, 2, 1]) == [0, 1]
    # same waste for a (4 pods on 2 nodes) and b (2 pods on 1 node): most-pods breaks the tie and the chain stops
    assert _both(oracle, enc, [LEAST_WASTE, MOST_PODS, LEAST_NODES], [2, 1, 1], [4, 2, 1], [4, 2, 1]) == [0]
    # most-pods ties a and b (4 pods each), least-nodes prefers b
    assert _both(oracle, enc, [MOST_PODS, LEAST_NODES], [2, 1, 1], [4, 4, 1], [4, 4, 1]) == [1]


# ---- price expander: expander/price/price_test.go:76-335 (TestPriceExpander), step by step -------------------------------
PRICE, PRIORITY = 3, 4


def _price_case(oracle, node_cpu, node_price, node_counts, option_pods, preferred_cpu, exists=None, price_error=None):
    """templates = the options' node groups; two single-pod groups p1 (20.0) and p2 (10.0), stabilization pod 10."""
    from kubernetes_autoscaler_b200.engine import expander_chain_ex
    p1, p2 = BuildTestPod("p1", 1000, 0), BuildTestPod("p2", 500, 0)
    groups = [makePodEquivalenceGroup(p1, 1), makePodEquivalenceGroup(p2, 1)]
    tmpl = [_node_info("n%d" % (i + 1), cpu, 1000) for i, cpu in enumerate(node_cpu)]
    enc = encode([], tmpl, groups)
    T = enc.T
    sched = np.asarray([[1 if g in pods else 0 for g in (0, 1)] for pods in option_pods], np.int32)
    order = np.asarray([[0, 1]] * T, np.int32)
    nc = np.asarray(node_counts, np.int32)
    pod_price = np.zeros(enc.struct.num_podspecs)
    pod_price[enc.arrays["pend_spec"][enc.arrays["group_off"][0]]] = 20.0
    pod_price[enc.arrays["pend_spec"][enc.arrays["group_off"][1]]] = 10.0
    score = oracle.price_scores(enc, node_price, pod_price, 10.0, preferred_cpu, exists=exists, node_count=nc, sched=sched, order=order)
    pc = sched.sum(axis=1)
    a = oracle.expander_ex([PRICE], nc, pc, price=score, price_error=price_error)
    b = expander_chain_ex([PRICE], nc, pc, price=score, price_error=price_error)
    assert np.array_equal(a, b)
    return [i for i in range(T) if a[i]], score


def test_price_expander_reference_sequence(oracle):
    both = [(0, 1), (0, 1)]
    # First node group is cheaper.
    assert _price_case(oracle, [1000, 4000], [20.0, 200.0], [2, 1], both, 2000)[0] == [0]
    # First node group is cheaper, however, the second one is preferred.
    assert _price_case(oracle, [1000, 4000], [50.0, 200.0], [2, 1], both, 4000)[0] == [1]
    # First node group is cheaper, the second is preferred but there is lots of nodes to be created.
    assert _price_case(oracle, [1000, 4000], [20.0, 200.0], [80, 40], both, 4000)[0] == [0]
    # Second node group is cheaper
    assert _price_case(oracle, [1000, 4000], [200.0, 100.0], [2, 1], both, 2000)[0] == [1]
    # First group accept 1 pod and second accepts 2: both equally expensive, however 2 accept two pods.
    assert _price_case(oracle, [1000, 4000], [200.0, 200.0], [2, 1], [(0,), (0, 1)], 2000)[0] == [1]
    # Errors are expected: no prices at all -> no option survives
    assert _price_case(oracle, [1000, 4000], [0.0, 0.0], [2, 1], [(0,), (0, 1)], 2000, price_error=[1, 1])[0] == []
    # Choose existing group when non-existing has the same price.
    three = [(0,), (0, 1), (0, 1)]
    assert _price_case(oracle, [1000, 4000, 4000], [200.0, 200.0, 200.0], [2, 1, 1], three, 2000, exists=[1, 1, 0])[0] == [1]
    # Choose non-existing group when non-existing is cheaper.
    assert _price_case(oracle, [1000, 4000, 4000], [200.0, 200.0, 90.0], [2, 1, 1], three, 2000, exists=[1, 1, 0])[0] == [2]


def test_price_score_formula_and_go_tanh(oracle):
    """The score of the first reference case by hand: ng1 = 2 nodes x 20 with p1 + p2, preferred 2000 m vs 1000 m."""
    import math
    _, score = _price_case(oracle, [1000, 4000], [20.0, 200.0], [2, 1], [(0, 1), (0, 1)], 2000)
    sub = (20.0 * 2 + 10.0) / ((0.0 + 20.0 + 10.0) + 10.0)
    supp = (2.0 - 1.0) * (1.0 - oracle.go_tanh(1.0 / 15.0)) + 1.0
    assert score[0] == supp * sub
    # the restated pure-Go math.Tanh agrees with libm to the last couple of ulps over the range the expander feeds it
    for n in list(range(1, 60)) + [100, 500, 1000, 5000]:
        x = (n - 1) / 15.0
        assert abs(oracle.go_tanh(x) - math.tanh(x)) <= 4 * math.ulp(1.0), n
    assert oracle.go_tanh(0.0) == 0.0 and oracle.go_tanh(50.0) == 1.0 and oracle.go_tanh(-50.0) == -1.0


# ---- priority expander: expander/priority/priority_test.go:107-135 ----------------------------------------------------------
_PRIO_CONFIG = {5: [r".*t2\.micro.*"], 10: [r".*t2\.large.*", r".*t3\.large.*"], 50: [r".*m4\.4xlarge.*"]}
_PRIO_NOT_MATCHING = {5: [r".*t\.micro.*"], 10: [r".*t\.large.*"]}
_PRIO_WILDCARD = {5: [r".*"], 10: [r".t2\.large.*"]}


@pytest.mark.parametrize("config,ids,want", [
    (_PRIO_CONFIG, ["my-asg.t2.large"], ["my-asg.t2.large"]),
    (_PRIO_CONFIG, ["my-asg.t2.large", "my-asg.m4.4xlarge"], ["my-asg.m4.4xlarge"]),
    (_PRIO_WILDCARD, ["my-asg.t2.large", "my-asg.t2.micro"], ["my-asg.t2.large"]),
    (_PRIO_CONFIG, ["my-asg.t2.large", "my-asg.t3.large", "my-asg.t2.micro"], ["my-asg.t2.large", "my-asg.t3.large"]),
    (_PRIO_NOT_MATCHING, ["my-asg.t2.large", "my-asg.t3.large"], ["my-asg.t2.large", "my-asg.t3.large"]),
], ids=["single-out-of-one", "single-out-of-many", "higher-priority-match", "two-out-of-many", "falls-back-to-all"])
def test_priority_expander_reference_cases(oracle, config, ids, want):
    from kubernetes_autoscaler_b200.engine import expander_chain_ex, resolve_priorities
    prio = resolve_priorities(config, ids)
    nc = np.ones(len(ids), np.int32)
    a = oracle.expander_ex([PRIORITY], nc, nc, priority=prio)
    b = expander_chain_ex([PRIORITY], nc, nc, priority=prio)
    assert np.array_equal(a, b)
    assert [ids[i] for i in range(len(ids)) if a[i]] == want


def test_chain_with_price_and_priority(oracle):
    """priority first, price among the survivors, then least-nodes (factory/chain.go:36-45)."""
    from kubernetes_autoscaler_b200.engine import expander_chain_ex
    nc, pc = np.asarray([2, 1, 1, 3], np.int32), np.asarray([4, 4, 4, 4], np.int32)
    prio = np.asarray([10, 10, 5, -1], np.int32)
    price = np.asarray([3.0, 3.0, 1.0, 0.5])
    for fn in (oracle.expander_ex, expander_chain_ex):
        assert fn([PRIORITY], nc, pc, priority=prio).tolist() == [1, 1, 0, 0]
        assert fn([PRIORITY, PRICE], nc, pc, price=price, priority=prio).tolist() == [1, 1, 0, 0]
        assert fn([PRIORITY, PRICE, LEAST_NODES], nc, pc, price=price, priority=prio).tolist() == [0, 1, 0, 0]
        assert fn([PRICE], nc, pc, price=price).tolist() == [0, 0, 0, 1]
