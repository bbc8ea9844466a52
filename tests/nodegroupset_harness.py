"""TEST HARNESS (not product code): host-side mirror of the step AFTER the path (first half): which similar node groups may share a
scale-up, and how the new nodes are split between them.  Small integer work on the host; its only input from the
engine is the exemplar feasibility matrix (``ScaleUpSimulation.schedulable_pod_groups``).

* ``matchingSchedulablePodGroups`` / ``ComputeSimilarNodeGroups``
  (``cluster-autoscaler/core/scaleup/orchestrator/orchestrator.go:757-812``).  The cluster-state safety check (``NodeGroupScaleUpSafety``) stays outside: the caller passes the candidate ids that
  passed it (``FindSimilarNodeGroups`` below gives the candidates).
* ``FindSimilarNodeGroups`` / ``IsCloudProviderNodeInfoSimilar`` (``processors/nodegroupset/balancing_processor.go:44-77``,
  ``compare_nodegroups.go:30-167``).
* ``BalanceScaleUpBetweenGroups`` (``processors/nodegroupset/balancing_processor.go:79-182``).  Go sorts the groups with the
  unstable ``sort.Slice``; ties keep the input order here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

from kubernetes_autoscaler_b200.estimator import NodeGroupInfo


@dataclass
class ScaleUpInfo:
    """processors/nodegroupset/nodegroup_set_processor.go:30-39."""
    group: NodeGroupInfo
    current_size: int
    new_size: int
    max_size: int


def matchingSchedulablePodGroups(podGroups: Sequence[int], similarPodGroups: Sequence[int]) -> bool:
    """orchestrator.go:800-812: every group the main node group can schedule is schedulable on the similar one too."""
    similar = set(similarPodGroups)
    return all(g in similar for g in podGroups)


def ComputeSimilarNodeGroups(node_group: str, similar_candidates: Sequence[str], schedulablePodGroups: Dict[str, List[int]],
                             balance_similar_node_groups: bool = True, zero_or_max_node_scaling: bool = False) -> List[str]:
    """orchestrator.go:757-798 with FindSimilarNodeGroups / NodeGroupScaleUpSafety already applied by the caller."""
    if not balance_similar_node_groups or zero_or_max_node_scaling:
        return []
    pod_groups = schedulablePodGroups.get(node_group)
    if not pod_groups:
        return []
    return [ng for ng in similar_candidates
            if ng in schedulablePodGroups and matchingSchedulablePodGroups(pod_groups, schedulablePodGroups[ng])]


def BalanceScaleUpBetweenGroups(groups: Sequence[NodeGroupInfo], newNodes: int) -> List[ScaleUpInfo]:
    """balancing_processor.go:79-182: nodes go to the smallest group first; MaxSize is respected; unchanged groups are
    dropped from the result."""
    if not groups:
        raise ValueError("Can't balance scale up between 0 groups")
    infos: List[ScaleUpInfo] = []
    total_capacity = 0
    for ng in groups:
        current, mx = ng.target_size, ng.max_size
        if current == mx:
            continue                                   # already maxed, ignore it
        if mx > current:
            total_capacity += mx - current
        infos.append(ScaleUpInfo(ng, current, current, mx))
    newNodes = min(newNodes, total_capacity)
    infos.sort(key=lambda i: i.current_size)           # stable
    start = cur = 0
    while newNodes > 0:
        info = infos[cur]
        if info.new_size < info.max_size:
            info.new_size += 1
            newNodes -= 1
        else:                                          # full (or over its max): swap it out of the active range
            infos[start], infos[cur] = infos[cur], infos[start]
            start += 1
        # Go holds a POINTER to the slot (currentInfo := &scaleUpInfos[currentIndex]): after the swap above it reads the
        # element that was swapped INTO the slot, not the full group that left it (balancing_processor.go:150-170)
        if cur < len(infos) - 1 and infos[cur].new_size > infos[cur + 1].new_size:
            cur += 1
        else:
            cur = start
    return [i for i in infos if i.new_size != i.current_size]


# ---- FindSimilarNodeGroups' comparator (processors/nodegroupset/compare_nodegroups.go:30-167) -----------------------------
BasicIgnoredLabels = {
    "kubernetes.io/hostname", "failure-domain.beta.kubernetes.io/zone", "failure-domain.beta.kubernetes.io/region",
    "topology.kubernetes.io/zone", "topology.kubernetes.io/region", "beta.kubernetes.io/fluentd-ds-ready",
    "kops.k8s.io/instancegroup",
}


@dataclass
class NodeGroupDifferenceRatios:
    """config/autoscaling_options.go:79-104 (NewDefaultNodeGroupDifferenceRatios)."""
    max_allocatable_difference_ratio: float = 0.05
    max_free_difference_ratio: float = 0.05
    max_capacity_memory_difference_ratio: float = 0.015


def _milli(name: str, v: int) -> float:
    # Quantity.MilliValue(): the object model keeps cpu in milli-cores already and everything else in base units
    return float(v) if name == "cpu" else float(v) * 1000.0


def _within_tolerance(name: str, values: List[int], ratio: float) -> bool:
    """resourceListWithinTolerance (:57-64)."""
    if len(values) != 2:
        return False
    a, b = _milli(name, values[0]), _milli(name, values[1])
    larger, smaller = max(a, b), min(a, b)
    return larger - smaller <= larger * ratio


def IsCloudProviderNodeInfoSimilar(n1, n2, ignoredLabels=frozenset(BasicIgnoredLabels),
                                   ratioOpts: NodeGroupDifferenceRatios = NodeGroupDifferenceRatios()) -> bool:
    """compare_nodegroups.go:104-163 on two NodeInfos of the object model (`requests` of the resident pods are the
    effective pod requests, as framework.NodeInfo.Requested sums them)."""
    capacity: Dict[str, List[int]] = {}
    allocatable: Dict[str, List[int]] = {}
    free: Dict[str, List[int]] = {}
    for ni in (n1, n2):
        for res, q in ni.node.capacity.items():
            capacity.setdefault(res, []).append(q)
        for res, q in ni.node.allocatable.items():
            allocatable.setdefault(res, []).append(q)
        requested: Dict[str, int] = {"cpu": 0, "memory": 0, "pods": 0, "ephemeral-storage": 0}   # ResourceToResourceList
        for p in ni.pods:
            for res, q in p.requests.items():
                requested[res] = requested.get(res, 0) + q
        for res, q in requested.items():
            free.setdefault(res, []).append(ni.node.allocatable.get(res, 0) - q)
    for kind, qty in capacity.items():
        if len(qty) != 2:
            return False                               # missing capacity
        if kind == "memory":
            if not _within_tolerance(kind, qty, ratioOpts.max_capacity_memory_difference_ratio):
                return False
        elif qty[0] != qty[1]:
            return False                               # every other capacity must match exactly
    if not all(_within_tolerance(k, v, ratioOpts.max_allocatable_difference_ratio) for k, v in allocatable.items()):
        return False
    if not all(_within_tolerance(k, v, ratioOpts.max_free_difference_ratio) for k, v in free.items()):
        return False
    labels: Dict[str, List[str]] = {}
    for ni in (n1, n2):
        for k, v in ni.node.labels.items():
            if k not in ignoredLabels:
                labels.setdefault(k, []).append(v)
    return all(len(v) == 2 and v[0] == v[1] for v in labels.values())


def CreateGenericNodeInfoComparator(extraIgnoredLabels: Sequence[str] = (), ratioOpts: NodeGroupDifferenceRatios = NodeGroupDifferenceRatios()):
    """compare_nodegroups.go:88-102."""
    ignored = frozenset(BasicIgnoredLabels) | frozenset(extraIgnoredLabels)
    return lambda a, b: IsCloudProviderNodeInfoSimilar(a, b, ignored, ratioOpts)


def FindSimilarNodeGroups(node_group: str, node_infos: Dict[str, object], comparator=None) -> List[str]:
    """balancing_processor.go:44-77: the other node groups whose template NodeInfo the comparator accepts."""
    comparator = comparator or CreateGenericNodeInfoComparator()
    if node_group not in node_infos:
        raise KeyError("failed to find template node for node group %s" % node_group)
    base = node_infos[node_group]
    return [ng for ng, ni in node_infos.items() if ng != node_group and comparator(base, ni)]
