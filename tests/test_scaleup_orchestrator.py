"""The scale-up tick composed on the host (kubernetes_autoscaler_b200/scaleup.py).  CPU part: the engine is replaced by a
test double that answers the same four calls from the oracle, so the orchestration logic is checked without a GPU; GPU part:
the same scenarios through the real engine must give the same plans."""
import numpy as np
import pytest

from kubernetes_autoscaler_b200.estimator import NodeGroupInfo
from kubernetes_autoscaler_b200.objects import BuildTestNode, BuildTestPod, NodeInfo, Taint, Toleration, WithTolerations
from scaleup_harness import (AutoscalingOptions, ScaleUpNoOptionsAvailable, ScaleUpOrchestrator, ScaleUpSuccessful)


class OracleEngine:
    """Test double with the Engine calls ScaleUpSimulation makes, answered by the CPU oracle."""

    def __init__(self):
        from oracle import pyoracle
        self.o = pyoracle

    def load(self, enc):
        self.enc = enc

    def feasibility_groups(self):
        return self.o.feasibility_groups(self.enc)

    def estimate_all(self, max_nodes=None, want_sched=True, copy=True):
        caps = None if max_nodes is None else np.asarray(max_nodes, np.int32)
        self.nc, self.pc, self.sched, order, _ = self.o.estimate_all(self.enc, caps)
        return self.nc, self.pc, self.sched, order

    def expander_best(self, chain, nc, pc, sched=None):
        return self.o.expander(self.enc, chain, nc, pc, self.sched if sched is None else sched)


def _pods(prefix, n, cpu, mem, uid, *opts):
    out = []
    for i in range(n):
        p = BuildTestPod("%s-%d" % (prefix, i), cpu, mem, *opts)
        p.owner_uid, p.owner_kind = uid, "ReplicaSet"
        out.append(p)
    return out


def _scenario():
    cluster = [NodeInfo(BuildTestNode("n%d" % i, 2000, 8 << 30), [BuildTestPod("r%d" % i, 1900, 1 << 30)]) for i in range(3)]
    small, big = BuildTestNode("small-t", 2000, 8 << 30), BuildTestNode("big-t", 8000, 32 << 30)
    gpu = BuildTestNode("gpu-t", 8000, 32 << 30)
    gpu.taints = [Taint("gpu", "true", "NoSchedule")]
    node_infos = {"small": NodeInfo(small), "small-b": NodeInfo(BuildTestNode("small-b-t", 2000, 8 << 30)), "big": NodeInfo(big), "gpu": NodeInfo(gpu)}
    groups = [NodeGroupInfo("small", 10, 1), NodeGroupInfo("small-b", 10, 3), NodeGroupInfo("big", 10, 0), NodeGroupInfo("gpu", 5, 5)]
    pods = (_pods("web", 12, 900, 1 << 30, "rs-web") + _pods("fat", 2, 6000, 16 << 30, "rs-fat") +
            _pods("cuda", 3, 1000, 1 << 30, "rs-cuda", WithTolerations(Toleration("gpu", "Equal", "true", "NoSchedule"))) +
            _pods("huge", 1, 64000, 1 << 30, "rs-huge"))
    return cluster, node_infos, groups, pods


def _check(status, chain):
    assert status.result == ScaleUpSuccessful
    names = lambda ps: sorted(p.name for p in ps)
    assert names(status.pods_remain_unschedulable) == ["huge-0"]            # fits no template
    plan = {i.group.id: (i.current_size, i.new_size) for i in status.scale_up_infos}
    if chain == ("least-nodes",):
        # big: 8000m / 32 GiB per node.  fat (6000m) first, web fills up: 14 pods + 3 cuda -> fewest nodes of all options
        assert list(plan) == ["big"] and plan["big"][0] == 0
        assert "fat-0" in names(status.pods_triggered_scale_up)
    return plan


@pytest.mark.parametrize("chain", [("least-waste",), ("most-pods", "least-nodes"), ("least-nodes",)])
def test_scale_up_on_the_oracle_double(chain):
    cluster, node_infos, groups, pods = _scenario()
    orch = ScaleUpOrchestrator(AutoscalingOptions(expander=chain), engine=OracleEngine())
    status = orch.ScaleUp(pods, cluster, node_infos, groups)
    plan = _check(status, chain)
    assert "gpu" not in plan                                                # gpu is at max size: skipped, never simulated
    assert sum(n - c for c, n in plan.values()) >= 1


def test_scale_up_balances_similar_groups_and_caps():
    cluster, node_infos, groups, pods = _scenario()
    web = [p for p in pods if p.name.startswith("web")]
    opts = AutoscalingOptions(expander=("least-waste",), balance_similar_node_groups=True, max_nodes_total=9)
    status = ScaleUpOrchestrator(opts, engine=OracleEngine()).ScaleUp(web, cluster, {k: node_infos[k] for k in ("small", "small-b")}, groups[:2])
    # 12 x 900m on 2000m nodes = 2 per node = 6 nodes; the cluster has 3 of at most 9 nodes: exactly 6 may be added
    assert status.result == ScaleUpSuccessful
    plan = {i.group.id: (i.current_size, i.new_size) for i in status.scale_up_infos}
    assert sum(n - c for c, n in plan.values()) == 6
    assert set(plan) == {"small", "small-b"} and plan["small"] == (1, 5) and plan["small-b"] == (3, 5)   # smallest group first
    opts.max_nodes_total = 5                                                # only 2 more nodes allowed
    status = ScaleUpOrchestrator(opts, engine=OracleEngine()).ScaleUp(web, cluster, {k: node_infos[k] for k in ("small", "small-b")}, groups[:2])
    assert sum(i.new_size - i.current_size for i in status.scale_up_infos) == 2
    assert len(status.pods_triggered_scale_up) == 4                         # the estimate was capped by the cluster-capacity threshold


def test_scale_up_no_options():
    cluster, node_infos, groups, pods = _scenario()
    huge = [p for p in pods if p.name.startswith("huge")]
    status = ScaleUpOrchestrator(engine=OracleEngine()).ScaleUp(huge, cluster, node_infos, groups)
    assert status.result == ScaleUpNoOptionsAvailable and [p.name for p in status.pods_remain_unschedulable] == ["huge-0"]
    status = ScaleUpOrchestrator(engine=OracleEngine()).ScaleUp(pods, cluster, node_infos, [NodeGroupInfo("gpu", 5, 5)])
    assert status.result == ScaleUpNoOptionsAvailable


@pytest.mark.gpu
@pytest.mark.parametrize("chain", [("least-waste",), ("most-pods", "least-nodes"), ("least-nodes",)])
def test_scale_up_engine_matches_the_oracle_double(chain):
    import __graft_entry__ as g
    g.build()
    from kubernetes_autoscaler_b200.engine import Engine
    cluster, node_infos, groups, pods = _scenario()
    eng = Engine(device=0)
    try:
        for balance in (False, True):
            opts = AutoscalingOptions(expander=chain, balance_similar_node_groups=balance, max_nodes_total=40)
            want = ScaleUpOrchestrator(opts, engine=OracleEngine()).ScaleUp(pods, cluster, node_infos, groups)
            got = ScaleUpOrchestrator(opts, engine=eng).ScaleUp(pods, cluster, node_infos, groups)
            assert got.result == want.result
            plan = lambda st: [(i.group.id, i.current_size, i.new_size) for i in st.scale_up_infos]
            assert plan(got) == plan(want)
            for f in ("pods_triggered_scale_up", "pods_remain_unschedulable", "pods_await_evaluation"):
                assert [p.name for p in getattr(got, f)] == [p.name for p in getattr(want, f)]
    finally:
        eng.close()
