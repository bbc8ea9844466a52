#!/usr/bin/env python
"""Writes tests/golden/snapshotz_small.json (a /snapshotz-format capture, v1.Node / v1.Pod JSON as the
apiserver serialises them) and tests/golden/snapshotz_small.expected.json (what the CPU oracle answers for
it).  Run from the repo root:  python tests/golden/make_snapshotz.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def node(name, cpu, mem, zone, pool, taints=(), pods="110", extra=None):
    alloc = {"cpu": cpu, "memory": mem, "pods": pods, "ephemeral-storage": "100Gi"}
    alloc.update(extra or {})
    return {"metadata": {"name": name, "labels": {"kubernetes.io/hostname": name, "topology.kubernetes.io/zone": zone,
                                                  "pool": pool, "node.kubernetes.io/instance-type": "m5.%s" % pool}},
            "spec": {"taints": [dict(key=k, value=v, effect=e) for k, v, e in taints]},
            "status": {"capacity": dict(alloc), "allocatable": dict(alloc)}}


def pod(name, ns, labels, containers, owner=None, **spec):
    meta = {"name": name, "namespace": ns, "labels": labels}
    if owner:
        meta["ownerReferences"] = [{"kind": "ReplicaSet", "name": owner, "uid": "uid-" + owner, "controller": True}]
    s = {"containers": [{"name": "c%d" % i, "resources": {"requests": r}, **({"ports": p} if p else {})}
                        for i, (r, p) in enumerate(containers)]}
    s.update(spec)
    return {"metadata": meta, "spec": s}


def main():
    ds = pod("kube-proxy-x", "kube-system", {"k8s-app": "kube-proxy"}, [({"cpu": "100m", "memory": "128Mi"}, None)],
             tolerations=[{"operator": "Exists"}])
    web_sel = {"matchLabels": {"app": "web"}}
    doc = {
        "NodeList": [
            {"Node": node("ip-10-0-0-1", "4", "16Gi", "us-east-1a", "general"),
             "Pods": [ds, pod("web-old-1", "prod", {"app": "web"}, [({"cpu": "500m", "memory": "1Gi"}, None)], owner="web-old")]},
            {"Node": node("ip-10-0-0-2", "4", "16Gi", "us-east-1b", "general"), "Pods": [ds]},
            {"Node": node("ip-10-0-0-3", "8", "32Gi", "us-east-1a", "gpu", taints=[("nvidia.com/gpu", "present", "NoSchedule")],
                          extra={"nvidia.com/gpu": "1"}),
             "Pods": [ds, pod("db-0", "prod", {"app": "db"}, [({"cpu": "2", "memory": "8Gi"}, None)],
                              affinity={"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
                                  {"labelSelector": {"matchLabels": {"app": "db"}}, "topologyKey": "kubernetes.io/hostname"}]}})]},
        ],
        "TemplateNodes": {
            "ng-general-1a": {"Node": node("template-general-1a", "3920m", "15Gi", "us-east-1a", "general"), "Pods": [ds]},
            "ng-general-1c": {"Node": node("template-general-1c", "7910m", "30.5Gi", "us-east-1c", "general"), "Pods": [ds]},
            "ng-gpu-1a": {"Node": node("template-gpu-1a", "15890m", "61Gi", "us-east-1a", "gpu",
                                       taints=[("nvidia.com/gpu", "present", "NoSchedule")], extra={"nvidia.com/gpu": "4"}), "Pods": [ds]},
            "ng-small-1b": {"Node": node("template-small-1b", "1930m", "3.5Gi", "us-east-1b", "small", pods="8"), "Pods": [ds]},
        },
        "UnscheduledPodsCanBeScheduled": [],
        "PendingPods": (
            [pod("web-%d" % i, "prod", {"app": "web"}, [({"cpu": "750m", "memory": "1536Mi"}, None), ({"cpu": "250m", "memory": "0.5Gi"}, None)],
                 owner="web", topologySpreadConstraints=[{"maxSkew": 1, "topologyKey": "topology.kubernetes.io/zone",
                                                          "whenUnsatisfiable": "DoNotSchedule", "labelSelector": web_sel}])
             for i in range(9)] +
            [pod("train-%d" % i, "ml", {"app": "train"}, [({"cpu": "3500m", "memory": "12Gi", "nvidia.com/gpu": "1"}, None)],
                 owner="train", tolerations=[{"key": "nvidia.com/gpu", "operator": "Exists", "effect": "NoSchedule"}],
                 nodeSelector={"pool": "gpu"}) for i in range(6)] +
            [pod("db-%d" % (i + 1), "prod", {"app": "db"}, [({"cpu": "2", "memory": "8Gi"}, None)], owner="db",
                 affinity={"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
                     {"labelSelector": {"matchLabels": {"app": "db"}}, "topologyKey": "kubernetes.io/hostname"}]}})
             for i in range(4)] +
            [pod("ingress-%d" % i, "edge", {"app": "ingress"}, [({"cpu": "200m", "memory": "256Mi"}, [{"containerPort": 8080, "hostPort": 443}])],
                 owner="ingress", initContainers=[{"name": "init", "resources": {"requests": {"cpu": "1", "memory": "64Mi"}}}])
             for i in range(5)] +
            [pod("batch-%d" % i, "batch", {"app": "batch"}, [({"cpu": "100m", "memory": "200M"}, None)], owner="batch",
                 overhead={"cpu": "50m", "memory": "10Mi"}) for i in range(40)]),
    }
    with open(os.path.join(HERE, "snapshotz_small.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)

    from kubernetes_autoscaler_b200.encode import encode
    from kubernetes_autoscaler_b200.snapshotz import load_snapshotz
    from oracle import pyoracle
    import numpy as np
    cluster, templates, groups, namespaces = load_snapshotz(doc)
    ids = list(templates.keys())
    enc = encode(cluster, [templates[i] for i in ids], groups, namespaces=namespaces)
    caps = np.full(enc.T, 10, np.int32)
    nc, pc, sched, order, _ = pyoracle.estimate_all(enc, caps)
    reasons = pyoracle.feasibility_groups(enc)
    exp = {"node_groups": ids, "groups": [[p.name for p in g.pods] for g in groups], "max_nodes": 10,
           "node_count": nc.tolist(), "pod_count": pc.tolist(), "sched_count": sched.tolist(),
           "order": order.tolist(), "group_reasons": reasons.tolist()}
    with open(os.path.join(HERE, "snapshotz_small.expected.json"), "w") as f:
        json.dump(exp, f, indent=1)
    print(json.dumps({k: exp[k] for k in ("node_groups", "node_count", "pod_count")}))


if __name__ == "__main__":
    main()
