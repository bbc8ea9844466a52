"""Host-side mirror of the scale-down consumer of the same primitive:
``RemovalSimulator.SimulateNodeRemoval`` / ``findPlaceFor`` (``cluster-autoscaler/simulator/cluster.go:126-217``).

``findPlaceFor`` is ``HintingSimulator.TrySchedulePods(snapshot without the node, its pods, isCandidateNode,
breakOnFailure=true)`` — exactly what ``cae_filter_schedulable`` runs on the GPU — so this file is bookkeeping only:
take the node out of the snapshot, clear ``spec.nodeName`` of the pods to move, ask the simulator, and (optionally)
persist a successful simulation into the snapshot.  Which pods must move (``GetPodsToMove``: drainability rules, PDBs)
is outside §8; the default here is every pod that is not DaemonSet-owned, or the caller passes the list.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from .engine import Engine
from .objects import Namespace, Node, NodeInfo, Pod
from .podlistprocessor import HintingSimulator

# simulator/cluster.go:55-95
NoReason, NoPlaceToMovePods, NoNodeInfo = "NoReason", "NoPlaceToMovePods", "NoNodeInfo"


@dataclass
class NodeToBeRemoved:
    """simulator/cluster.go:41-51."""
    node: Node
    pods_to_reschedule: List[Pod] = field(default_factory=list)
    daemonset_pods: List[Pod] = field(default_factory=list)


@dataclass
class UnremovableNode:
    """simulator/cluster.go:53-58."""
    node: Node
    reason: str


class RemovalSimulator:
    def __init__(self, cluster_snapshot: List[NodeInfo], persistSuccessfulSimulations: bool = False,
                 engine: Optional[Engine] = None, schedulingSimulator: Optional[HintingSimulator] = None) -> None:
        self.cluster = cluster_snapshot          # mutated only when a successful simulation is persisted
        self.canPersist = persistSuccessfulSimulations
        self.schedulingSimulator = schedulingSimulator or HintingSimulator(engine)

    def SimulateNodeRemoval(self, nodeName: str, destinationMap: Dict[str, bool], pods_to_move: Optional[Sequence[Pod]] = None,
                            namespaces: Sequence[Namespace] = ()) -> Tuple[Optional[NodeToBeRemoved], Optional[UnremovableNode]]:
        """Exactly one of the two results is set (cluster.go:126-167)."""
        ni = next((n for n in self.cluster if n.node.name == nodeName), None)
        if ni is None:
            return None, UnremovableNode(Node(name=nodeName), NoNodeInfo)
        daemonset = [p for p in ni.pods if p.owner_kind == "DaemonSet"]
        to_move = list(pods_to_move) if pods_to_move is not None else [p for p in ni.pods if p.owner_kind != "DaemonSet"]
        placements = self._findPlaceFor(ni, to_move, destinationMap, namespaces)
        if placements is None:
            return None, UnremovableNode(ni.node, NoPlaceToMovePods)
        if self.canPersist:                       # withForkedSnapshot: Commit (cluster.go:169-182)
            self.cluster.remove(ni)
            by_name = {n.node.name: n for n in self.cluster}
            for pod, node_name in placements:
                by_name[node_name].pods.append(pod)
        return NodeToBeRemoved(ni.node, to_move, daemonset), None

    def _findPlaceFor(self, removed: NodeInfo, pods: Sequence[Pod], nodes: Dict[str, bool], namespaces):
        """cluster.go:184-217: the node leaves the snapshot first so that it does not take part in topology spreading."""
        removed_name = removed.node.name
        snapshot = [n for n in self.cluster if n is not removed]
        newpods = []
        for p in pods:
            q = p.clone()
            q.node_name = ""
            newpods.append(q)
        if not newpods:
            return []
        statuses, _ = self.schedulingSimulator.TrySchedulePods(
            snapshot, newpods, lambda ni: ni.node.name != removed_name and bool(nodes.get(ni.node.name)), True, namespaces)
        if len(statuses) != len(newpods):
            return None                           # "can reschedule only %d out of %d pods"
        return [(s.pod, s.node_name) for s in statuses]

    def DropOldHints(self) -> None:
        self.schedulingSimulator.DropOldHints()


NewRemovalSimulator = RemovalSimulator
