"""Host-side edge helpers: PodRequests and BuildPodGroups against the reference's examples."""
from kubernetes_autoscaler_b200.objects import BuildTestPod
from kubernetes_autoscaler_b200.podutil import Container, build_pod_groups, pod_requests

G = 1_000_000_000


def test_pod_requests_doc_example():
    """K8S/framework/plugins/noderesources/fit.go:266-292: IC1 2cpu/1G, IC2 2cpu/3G, C1 2cpu/1G, C2 1cpu/1G -> 3 cpu / 3G."""
    r = pod_requests([Container({"cpu": 2000, "memory": 1 * G}), Container({"cpu": 1000, "memory": 1 * G})],
                     [Container({"cpu": 2000, "memory": 1 * G}), Container({"cpu": 2000, "memory": 3 * G})])
    assert r == {"cpu": 3000, "memory": 3 * G}


def test_pod_requests_sidecars_overhead_podlevel():
    # a restartable init container counts for the whole lifetime and on top of later init containers
    r = pod_requests([Container({"cpu": 100})],
                     [Container({"cpu": 50}, restart_policy_always=True), Container({"cpu": 500})],
                     overhead={"cpu": 10})
    assert r["cpu"] == max(100 + 50, 500 + 50) + 10
    # pod-level requests replace the aggregate for cpu / memory only (PodLevelResources on)
    r = pod_requests([Container({"cpu": 100, "memory": 5, "nvidia.com/gpu": 1})], pod_level={"cpu": 700, "nvidia.com/gpu": 9})
    assert r == {"cpu": 700, "memory": 5, "nvidia.com/gpu": 1}


def test_build_pod_groups_reference_case():
    """core/scaleup/equivalence/groups_test.go:70-138 (the volume-based p4/p5 cases need volumes, which the
    engine refuses anyway): ownerless pod alone, pods of one controller with equal spec together."""
    def owned(name, cpu, uid):
        p = BuildTestPod(name, cpu, 200000)
        p.owner_uid, p.owner_kind = uid, "ReplicationController"
        return p
    p1 = BuildTestPod("p1", 1500, 200000)
    p2_1, p2_2 = owned("p2_1", 3000, "rc1"), owned("p2_2", 3000, "rc1")
    p3_1, p3_2 = owned("p3_1", 100, "rc2"), owned("p3_2", 100, "rc2")
    groups = build_pod_groups([p1, p2_1, p2_2, p3_1, p3_2])
    got = sorted(sorted(p.name for p in g.pods) for g in groups)
    assert got == [["p1"], ["p2_1", "p2_2"], ["p3_1", "p3_2"]]


def test_build_pod_groups_limit_per_controller():
    """groups.go:58,80-88: at most 10 equivalence groups per controller; later distinct pods become singletons
    that are not remembered."""
    pods = []
    for i in range(12):
        for rep in range(2):
            p = BuildTestPod("p%d_%d" % (i, rep), 100 + i, 1000)
            p.owner_uid, p.owner_kind = "rc", "ReplicaSet"
            pods.append(p)
    groups = build_pod_groups(pods)
    sizes = sorted(len(g.pods) for g in groups)
    assert sizes == [1, 1, 1, 1] + [2] * 10
    ds = BuildTestPod("ds", 1, 1)
    ds.owner_uid, ds.owner_kind = "d", "DaemonSet"
    assert [len(g.pods) for g in build_pod_groups([ds, ds])] == [1, 1]


def test_build_pod_groups_ignores_daemonsets():
    """core/scaleup/equivalence/groups_test.go:170-187: DaemonSet-owned pods are never grouped."""
    from kubernetes_autoscaler_b200.objects import BuildTestPod
    from kubernetes_autoscaler_b200.podutil import build_pod_groups
    pods = [BuildTestPod("p1", 3000, 200000), BuildTestPod("p2", 3000, 200000)]
    for p in pods:
        p.owner_uid, p.owner_kind = "12345678-1234-1234-1234-123456789012", "DaemonSet"
    assert len(build_pod_groups(pods)) == 2
