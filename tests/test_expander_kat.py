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
