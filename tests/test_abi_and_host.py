"""CPU-side checks: the C-ABI library loads and exports every symbol include/caengine.h declares,
the ctypes structs match the header, the encoder builds consistent tables, and the host-side limiter
logic (product code, Python) agrees with the reference tables and the oracle."""
import ctypes
import os

import numpy as np
import pytest

from kubernetes_autoscaler_b200 import capi, synth
from kubernetes_autoscaler_b200.encode import encode
from kubernetes_autoscaler_b200.estimator import (ClusterCapacityThreshold, EstimationContext, NodeGroupInfo,
                                                  SngCapacityThreshold, StaticThreshold,
                                                  ThresholdBasedEstimationLimiter, getMinLimit)
from kubernetes_autoscaler_b200.objects import (BuildTestNode, BuildTestPod, LabelSelector, NodeInfo,
                                                PodAffinityTerm, Requirement, NodeSelectorTerm, Toleration,
                                                WithLabels, makePodEquivalenceGroup)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(capi.ENGINE_LIB)
    names = capi.declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_version_string():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(capi.ENGINE_LIB)
    lib.cae_version.restype = ctypes.c_char_p
    assert b"caengine" in lib.cae_version()


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: without a device cae_create must return an error, not a dummy engine."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kubernetes_autoscaler_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine()


def test_struct_layout_matches_header():
    # every pointer field of cae_objects is filled by the encoder; sizes are 8-byte aligned
    enc = synth.generate(1)
    assert ctypes.sizeof(capi.cae_objects) % 8 == 0
    assert enc.struct.abi_version == capi.CONST["CAE_ABI_VERSION"]
    assert enc.P == 1000 and enc.T == 50
    assert enc.arrays["group_off"][-1] == enc.P


def test_encoder_tables():
    pod = BuildTestPod("p", 100, 200, WithLabels({"app": "x"}))
    pod.tolerations = [Toleration(key="k", operator="Exists", effect="NoSchedule")]
    pod.node_selector = {"pool": "a"}
    pod.node_affinity_terms = [NodeSelectorTerm(match_expressions=[Requirement("zone", "In", ["z1", "z2"])]),
                               NodeSelectorTerm()]
    pod.pod_anti_affinity = [PodAffinityTerm(LabelSelector(match_labels={"app": "x"}), "kubernetes.io/hostname")]
    node = BuildTestNode("n1", 1000, 1000)
    node.labels = {"pool": "a", "zone": "z1", "kubernetes.io/hostname": "n1"}
    enc = encode([NodeInfo(node)], [NodeInfo(BuildTestNode("t", 2000, 2000))], [makePodEquivalenceGroup(pod, 3)])
    s = enc.struct
    assert s.num_cluster_nodes == 1 and s.num_templates == 1 and s.num_groups == 1 and s.num_pending == 3
    assert s.num_naff == 1 and s.num_naff_terms == 2          # the empty term is kept
    assert s.hostname_key >= 0 and s.unschedulable_taint_key >= 0
    a = enc.arrays
    spec = a["pend_spec"][0]
    assert a["ps_req"][spec][0] == 100 and a["ps_req"][spec][1] == 200
    anti = a["ps_anti_list"][spec]
    t = a["aff_off"][anti]
    assert a["aterm_ns"][a["aterm_ns_off"][t]] == a["ps_namespace"][spec]  # own namespace defaulted in
    assert np.all(a["pend_spec"] == spec)


def test_get_min_limit_matches_reference_table():
    # threshold_based_limiter_test.go:187-203
    for base, target, want in [(-10, 10, -1), (-10, 0, -1), (-10, -10, -1), (0, 0, 0), (0, 10, 10), (5, 10, 5)]:
        assert getMinLimit(base, target) == want


def test_python_thresholds_match_oracle(oracle):
    ctx = EstimationContext(similar_node_groups=[NodeGroupInfo("a", 10, 5), NodeGroupInfo("b", 100, 50), NodeGroupInfo("c", 5, 3)],
                            cluster_max_node_limit=10, current_node_count=5)
    main = NodeGroupInfo("main", 20, 10)
    assert SngCapacityThreshold().NodeLimit(main, ctx) == 67 == oracle.sng_capacity_limit(True, [20, 10, 100, 5], [10, 5, 50, 3])
    assert ClusterCapacityThreshold().NodeLimit(main, ctx) == 5 == oracle.cluster_capacity_limit(True, 10, 5)
    assert ClusterCapacityThreshold().NodeLimit(main, None) == 0
    lim = ThresholdBasedEstimationLimiter([StaticThreshold(1000), SngCapacityThreshold(), ClusterCapacityThreshold()])
    assert lim.max_nodes(main, ctx) == 5
    assert ThresholdBasedEstimationLimiter([StaticThreshold(-1), StaticThreshold(10)]).max_nodes() == -1
    assert ThresholdBasedEstimationLimiter([]).max_nodes() == 0


def test_synth_is_deterministic():
    a, b = synth.generate(1), synth.generate(1)
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_zero_or_max_node_scaling():
    """core/scaleup/orchestrator/orchestrator.go:505-517 (TestZeroOrMaxNodeScaling, orchestrator_test.go:133)."""
    from kubernetes_autoscaler_b200.estimator import apply_zero_or_max
    pods = [BuildTestPod("p", 1, 1)]
    assert apply_zero_or_max(3, pods, 10, False) == (10, pods)      # raised to the only valid size
    assert apply_zero_or_max(12, pods, 10, False) == (10, pods)     # capped
    assert apply_zero_or_max(12, pods, 10, True) == (0, [])         # all-or-nothing refuses to cap
    assert apply_zero_or_max(0, [], 10, False) == (0, [])
