"""Host-side mirror of the step AFTER the path (first half): which similar node groups may share a
scale-up, and how the new nodes are split between them.  Small integer work on the host; its only input from the
engine is the exemplar feasibility matrix (``ScaleUpSimulation.schedulable_pod_groups``).

* ``matchingSchedulablePodGroups`` / ``ComputeSimilarNodeGroups``
  (``cluster-autoscaler/core/scaleup/orchestrator/orchestrator.go:757-812``).  ``FindSimilarNodeGroups`` (label / capacity
  comparator, ``processors/nodegroupset/compare_nodegroups.go``) and the cluster-state safety check stay outside:
  the caller passes the candidate ids that passed them.
* ``BalanceScaleUpBetweenGroups`` (``processors/nodegroupset/balancing_processor.go:79-182``).  Go sorts the groups with the
  unstable ``sort.Slice``; ties keep the input order here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

from .estimator import NodeGroupInfo


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
        if cur < len(infos) - 1 and info.new_size > infos[cur + 1].new_size:
            cur += 1
        else:
            cur = start
    return [i for i in infos if i.new_size != i.current_size]
