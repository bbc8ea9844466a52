"""/snapshotz ingest (real v1.Node / v1.Pod JSON -> object model -> tables) against the committed golden
fixture tests/golden/snapshotz_small.json (+ .expected.json, both written by tests/golden/make_snapshotz.py)."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from kubernetes_autoscaler_b200.encode import encode
from kubernetes_autoscaler_b200.snapshotz import (load_snapshotz, parse_quantity, pod_from_json, quantity_milli,
                                                  quantity_value)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    doc = json.load(open(os.path.join(GOLD, "snapshotz_small.json")))
    exp = json.load(open(os.path.join(GOLD, "snapshotz_small.expected.json")))
    cluster, templates, groups, namespaces = load_snapshotz(doc)
    ids = list(templates.keys())
    assert ids == exp["node_groups"]
    assert [[p.name for p in g.pods] for g in groups] == exp["groups"]
    enc = encode(cluster, [templates[i] for i in ids], groups, namespaces=namespaces)
    return enc, exp


def test_quantities():
    """resource.Quantity forms and the round-up of Value()/MilliValue()."""
    assert parse_quantity("100m") == Fraction(1, 10) and quantity_milli("100m") == 100
    assert quantity_value("1Gi") == 1 << 30 and quantity_value("1.5Gi") == 3 << 29
    assert quantity_value("200M") == 200_000_000 and quantity_value("1e3") == 1000
    assert quantity_milli("2") == 2000 and quantity_milli("0.5") == 500 and quantity_milli("1500m") == 1500
    assert quantity_value("100m") == 1            # Value() rounds up
    assert quantity_milli("1n") == 1              # MilliValue() rounds up
    assert quantity_value("30.5Gi") == 32749125632
    with pytest.raises(ValueError):
        parse_quantity("12 apples")


def test_pod_from_json_requests_ports_owner():
    p = pod_from_json({
        "metadata": {"name": "x", "namespace": "n", "labels": {"a": "b"},
                     "ownerReferences": [{"kind": "ReplicaSet", "uid": "u1", "controller": True}]},
        "spec": {"containers": [{"resources": {"requests": {"cpu": "250m", "memory": "64Mi"}}, "ports": [{"hostPort": 80}]},
                                {"resources": {"requests": {"cpu": "1"}}, "ports": [{"containerPort": 1}]}],
                 "initContainers": [{"resources": {"requests": {"cpu": "2"}}},
                                    {"restartPolicy": "Always", "resources": {"requests": {"memory": "1Mi"}}, "ports": [{"hostPort": 53, "protocol": "UDP"}]}],
                 "overhead": {"cpu": "10m"}, "volumes": [{"name": "v", "emptyDir": {}}]}})
    assert p.requests["cpu"] == 2000 + 10 and p.requests["memory"] == (64 << 20) + (1 << 20)
    assert [(h.host_port, h.protocol) for h in p.host_ports] == [(80, ""), (53, "UDP")]
    assert (p.owner_uid, p.owner_kind) == ("u1", "ReplicaSet") and not p.has_volumes_or_claims
    q = pod_from_json({"metadata": {"name": "y"}, "spec": {"containers": [], "volumes": [{"persistentVolumeClaim": {"claimName": "c"}}]}})
    assert q.has_volumes_or_claims


def test_golden_fixture_oracle(oracle):
    enc, exp = _load()
    caps = np.full(enc.T, exp["max_nodes"], np.int32)
    nc, pc, sched, order, _ = oracle.estimate_all(enc, caps)
    assert nc.tolist() == exp["node_count"] and pc.tolist() == exp["pod_count"]
    assert sched.tolist() == exp["sched_count"] and order.tolist() == exp["order"]
    assert oracle.feasibility_groups(enc).tolist() == exp["group_reasons"]


@pytest.mark.gpu
def test_golden_fixture_engine():
    """The engine reproduces the committed golden answers (no oracle involved)."""
    import __graft_entry__ as g
    g.build()
    from kubernetes_autoscaler_b200.engine import Engine
    enc, exp = _load()
    eng = Engine()
    eng.load(enc)
    assert eng.feasibility_groups().tolist() == exp["group_reasons"]
    nc, pc, sched, order = eng.estimate_all(np.full(enc.T, exp["max_nodes"], np.int32))
    assert nc.tolist() == exp["node_count"] and pc.tolist() == exp["pod_count"]
    assert sched.tolist() == exp["sched_count"] and order.tolist() == exp["order"]
    eng.close()


def test_volumes_refuse_pending_pods_only():
    """A RESIDENT pod with a PVC cannot make a volume filter reject anybody (they only look at the incoming pod's volumes):
    the tick stays on the engine; a PENDING pod with volumes is handed to the stock path (encode.Unsupported)."""
    from kubernetes_autoscaler_b200.encode import Unsupported, encode
    from kubernetes_autoscaler_b200.objects import BuildTestNode, BuildTestPod, NodeInfo, makePodEquivalenceGroup
    resident = BuildTestPod("db-0", 100, 1 << 20)
    resident.has_volumes_or_claims = True
    cluster = [NodeInfo(BuildTestNode("n1", 1000, 1 << 30), [resident])]
    templates = [NodeInfo(BuildTestNode("t1", 1000, 1 << 30))]
    enc = encode(cluster, templates, [makePodEquivalenceGroup(BuildTestPod("web", 100, 1 << 20), 3)])
    assert enc.P == 3
    pending = BuildTestPod("db-1", 100, 1 << 20)
    pending.has_volumes_or_claims = True
    with pytest.raises(Unsupported):
        encode(cluster, templates, [makePodEquivalenceGroup(pending, 1)])
    from kubernetes_autoscaler_b200.snapshotz import pod_from_json
    p = pod_from_json({"metadata": {"name": "x", "namespace": "d"}, "spec": {"volumes": [{"name": "v", "gcePersistentDisk": {"pdName": "d"}}], "containers": []}})
    assert p.has_volumes_or_claims
    p = pod_from_json({"metadata": {"name": "x", "namespace": "d"}, "spec": {"volumes": [{"name": "v", "configMap": {"name": "c"}}], "containers": []}})
    assert not p.has_volumes_or_claims
