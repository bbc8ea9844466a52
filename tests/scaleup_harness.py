"""TEST HARNESS (not product code): the Go side keeps ScaleUpOrchestrator.ScaleUp; this mirror exists
so that the engine's answers can be pinned on the reference's decision-level tests (orchestrator_test.go).

The scale-up tick end to end on the host, composed from the pieces the engine accelerates — the reading order of
``ScaleUpOrchestrator.ScaleUp`` (``cluster-autoscaler/core/scaleup/orchestrator/orchestrator.go:87-285``):

    BuildPodGroups -> valid node groups -> SchedulablePodGroups (engine, E x T) -> limiter caps -> Estimate for every
    node group (engine, one pass) -> options -> expander chain (engine) -> GetCappedNewNodeCount -> similar node groups
    -> BalanceScaleUpBetweenGroups -> scale-up plan.

Deliberately left to the Go side (outside SURVEY §8): resource quotas (``applyLimits``), node-group creation
(``CreateNodeGroup*``), the cluster-state registry (every group counts as ready), executing the plan, and the random
tie-break of the expander chain (the first surviving option is taken, so the result is deterministic).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

from kubernetes_autoscaler_b200.engine import Engine
from kubernetes_autoscaler_b200.estimator import (EstimationContext, NodeGroupInfo, Option, ScaleUpSimulation, ThresholdBasedEstimationLimiter,
                        ClusterCapacityThreshold, SngCapacityThreshold, StaticThreshold)
from nodegroupset_harness import (BalanceScaleUpBetweenGroups, ComputeSimilarNodeGroups, CreateGenericNodeInfoComparator,
                           FindSimilarNodeGroups, ScaleUpInfo)
from kubernetes_autoscaler_b200.objects import Namespace, NodeInfo, Pod
from kubernetes_autoscaler_b200.podutil import build_pod_groups

ScaleUpSuccessful, ScaleUpNoOptionsAvailable, ScaleUpError = "ScaleUpSuccessful", "ScaleUpNoOptionsAvailable", "ScaleUpError"


@dataclass
class AutoscalingOptions:
    """The fields of config.AutoscalingOptions this path reads."""
    max_nodes_total: int = 0                       # --max-nodes-total (0 = unlimited)
    max_nodes_per_scaleup: int = 1000              # --max-nodes-per-scaleup (static threshold)
    balance_similar_node_groups: bool = False      # --balance-similar-node-groups
    expander: Sequence[str] = ("least-waste",)     # --expander chain (random is the implicit fallback)
    ignored_labels: Sequence[str] = ()             # --balancing-ignore-label


@dataclass
class ScaleUpStatus:
    """processors/status/scale_up_status_processor.go:31-47."""
    result: str
    scale_up_infos: List[ScaleUpInfo] = field(default_factory=list)
    pods_triggered_scale_up: List[Pod] = field(default_factory=list)
    pods_remain_unschedulable: List[Pod] = field(default_factory=list)
    pods_await_evaluation: List[Pod] = field(default_factory=list)
    considered_node_groups: List[str] = field(default_factory=list)
    error: str = ""


class ScaleUpOrchestrator:
    def __init__(self, options: Optional[AutoscalingOptions] = None, engine: Optional[Engine] = None) -> None:
        self.options = options or AutoscalingOptions()
        self.engine = engine

    def GetCappedNewNodeCount(self, newNodeCount: int, currentNodeCount: int) -> int:
        """orchestrator.go:716-729; raises when the cluster is already at --max-nodes-total."""
        mx = self.options.max_nodes_total
        if mx > 0 and newNodeCount + currentNodeCount > mx:
            newNodeCount = mx - currentNodeCount
            if newNodeCount < 1:
                raise RuntimeError("max node total count already reached")
        return newNodeCount

    def ScaleUp(self, unschedulablePods: Sequence[Pod], cluster: Sequence[NodeInfo], nodeInfos: Dict[str, NodeInfo],
                nodeGroups: Sequence[NodeGroupInfo], allOrNothing: bool = False,
                namespaces: Sequence[Namespace] = (), simulation_factory=ScaleUpSimulation,
                expander_strategy=None, binpacking_limiter=None) -> ScaleUpStatus:
        """expander_strategy(options) -> Option: stands for expander.Strategy.BestOption when given (the reference tests use a
        reporting mock); binpacking_limiter: object with StopBinpacking(options) -> bool, asked after every node group in
        order (processors/binpacking/binpacking_limiter.go) — the engine computes every group in one pass, the limiter cuts
        the option list where the sequential loop would have stopped."""
        groups = build_pod_groups(list(unschedulablePods))                                   # :107
        considered = [ng.id for ng in nodeGroups]
        by_id = {ng.id: ng for ng in nodeGroups}
        # filterValidScaleUpNodeGroups (:417-460), the part that needs no cloud provider: max size reached, no template
        valid = [ng for ng in nodeGroups if ng.id in nodeInfos and ng.target_size < ng.max_size]
        all_pods = [p for g in groups for p in g.pods]
        if not valid or not groups:
            return ScaleUpStatus(ScaleUpNoOptionsAvailable, pods_remain_unschedulable=all_pods, considered_node_groups=considered)
        templates = {ng.id: nodeInfos[ng.id] for ng in valid}
        sim = simulation_factory(cluster, templates, groups, self.engine, namespaces)
        schedulable = sim.schedulable_pod_groups()                                           # :144-146
        # limiter caps per node group (estimator.NewDefaultEstimationLimiter: static, cluster capacity, similar groups)
        comparator = CreateGenericNodeInfoComparator(self.options.ignored_labels)
        similar: Dict[str, List[str]] = {}
        caps: Dict[str, int] = {}
        limiter = ThresholdBasedEstimationLimiter([StaticThreshold(self.options.max_nodes_per_scaleup), ClusterCapacityThreshold(),
                                                   SngCapacityThreshold()])
        for ng in valid:
            cands = FindSimilarNodeGroups(ng.id, templates, comparator) if self.options.balance_similar_node_groups else []
            similar[ng.id] = ComputeSimilarNodeGroups(ng.id, cands, schedulable, self.options.balance_similar_node_groups)
            ctx = EstimationContext([by_id[s] for s in similar[ng.id]], self.options.max_nodes_total, len(cluster))
            caps[ng.id] = limiter.max_nodes(ng, ctx)
        options: List[Option] = []
        for opt in sim.compute_expansion_options(caps):                                      # :148-162
            if allOrNothing and len(opt.pods) < len(all_pods):
                continue
            if opt.node_count > 0 and opt.pods:
                options.append(opt)
            if binpacking_limiter is not None and binpacking_limiter.StopBinpacking(options):   # :164-166
                break
        self.last_options = list(options)
        schedulable_somewhere = {g for gs in schedulable.values() for g in gs}
        # GetRemainingPods (:843-857): pods of the groups that are schedulable on NO node group
        remain = [p for gi, g in enumerate(groups) if gi not in schedulable_somewhere for p in g.pods]
        if not options:
            return ScaleUpStatus(ScaleUpNoOptionsAvailable, pods_remain_unschedulable=remain, considered_node_groups=considered)
        if expander_strategy is not None:
            best = expander_strategy(options)                                                # :178
            if best is None:
                return ScaleUpStatus(ScaleUpNoOptionsAvailable, pods_remain_unschedulable=remain, considered_node_groups=considered)
        else:
            surviving = sim.best_options(list(self.options.expander), options)               # only what reached the expander
            if not surviving:
                return ScaleUpStatus(ScaleUpNoOptionsAvailable, pods_remain_unschedulable=remain, considered_node_groups=considered)
            best = next(o for o in options if o.node_group == surviving[0])
        try:
            newNodes = self.GetCappedNewNodeCount(best.node_count, len(cluster))             # :194
        except RuntimeError as ex:
            return ScaleUpStatus(ScaleUpError, pods_triggered_scale_up=best.pods, considered_node_groups=considered, error=str(ex))
        targets = [by_id[best.node_group]] + [by_id[s] for s in similar[best.node_group]]    # balanceScaleUps (:731-755)
        infos = BalanceScaleUpBetweenGroups(targets, newNodes)
        if sum(i.new_size - i.current_size for i in infos) < newNodes and allOrNothing:      # :248-258
            return ScaleUpStatus(ScaleUpNoOptionsAvailable, pods_remain_unschedulable=all_pods, considered_node_groups=considered)
        # GetPodsAwaitingEvaluation (:859-870): schedulable on some node group, but not on the one that was picked
        on_best = set(schedulable[best.node_group])
        await_eval = [p for gi, g in enumerate(groups) if gi in schedulable_somewhere and gi not in on_best for p in g.pods]
        return ScaleUpStatus(ScaleUpSuccessful, infos, best.pods, remain, await_eval, considered)
