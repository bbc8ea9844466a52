"""Parity at the REAL sizes of BASELINE.json's configurations (the miniatures live in test_gpu_parity.py):

  C2  100 000 pods x 1 000 templates: every one of the 10^8 cells of the dense pass vs the oracle
  C3  100 000 x 5 000 (+PodTopologySpread), C4 500 000 x 5 000 (+InterPodAffinity), C5 1 000 000 x 10 000:
      the engine runs the FULL configuration; the oracle (whose PreFilter rescans the cluster per evaluation,
      ~0.3 ms per cell on C5) checks a slice: templates spread over the range x pod chunks spread over the range for
      the dense pass, every group exemplar and the whole Estimate() (node count, pod count, per-group scheduled
      counts, processing order) on the same templates.  Bit-exact bar (integer / index work).
The oracle runs on a pool of host processes (tests/oracle_pool.py)."""
import numpy as np
import pytest

from kubernetes_autoscaler_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    g.build()
    from kubernetes_autoscaler_b200.engine import Engine
    e = Engine(device=0, want_reasons=False)
    yield e
    e.close()


def _spread(n, k):
    return sorted({int(round(i * (n - 1) / max(k - 1, 1))) for i in range(k)})


def test_c2_full_dense_every_cell(eng, oracle):
    """BASELINE config 2 at size: all 10^8 (pod, template) verdicts and the fit histogram."""
    from oracle_pool import OraclePool
    from kubernetes_autoscaler_b200.engine import unpack_bits
    enc = synth.generate(2)
    eng.load(enc)
    bits, _, count = eng.feasibility()
    fit = unpack_bits(bits, enc.P)
    with OraclePool(2) as pool:
        cuts = [enc.T * i // (4 * pool.procs) for i in range(4 * pool.procs + 1)]
        jobs = [((0, enc.P), (cuts[i], cuts[i + 1])) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]
        cells = 0
        for (tb, te), _, reasons in pool.dense(jobs):
            want = reasons == 0
            assert np.array_equal(fit[tb:te], want), "dense verdicts differ in templates %d..%d" % (tb, te)
            assert np.array_equal(count[tb:te], want.sum(axis=1))
            cells += want.size
    assert cells == enc.P * enc.T == 100_000_000


@pytest.mark.parametrize("config,n_templates,n_chunks", [(3, 16, 6), (4, 8, 4), (5, 6, 3)], ids=["C3", "C4", "C5"])
def test_full_config_slices(eng, oracle, config, n_templates, n_chunks):
    from oracle_pool import OraclePool
    from kubernetes_autoscaler_b200.engine import unpack_bits
    enc = synth.generate(config)
    cfg = synth.CONFIGS[config]
    assert (enc.P, enc.T) == (cfg.pods, cfg.templates)
    eng.load(enc)
    bits, _, count = eng.feasibility()
    greasons = eng.feasibility_groups()
    cap = 1000
    nc, pc, sched, order = eng.estimate_all(np.full(enc.T, cap, np.int32))
    templates = _spread(enc.T, n_templates)
    chunk = 192
    starts = [s // 32 * 32 for s in _spread(enc.P - chunk, n_chunks)]
    with OraclePool(config) as pool:
        # dense verdicts on (template, pod chunk) samples
        jobs = [((s, s + chunk), (t, t + 1)) for t in templates for s in starts]
        for (tb, _), (pb, pe), reasons in pool.dense(jobs):
            got = unpack_bits(bits[tb:tb + 1, pb // 32:(pe + 31) // 32], pe - pb)
            assert np.array_equal(got, reasons == 0), "dense verdicts differ: template %d pods %d..%d" % (tb, pb, pe)
        # every group exemplar on the slice templates (SchedulablePodGroups)
        for t, want in pool.groups(templates).items():
            assert np.array_equal(greasons[t], want), "group reasons differ on template %d" % t
        # the whole Estimate() on the slice templates
        for t, (onc, opc, osched, oorder) in pool.estimate(templates, cap).items():
            assert (int(nc[t]), int(pc[t])) == (onc, opc), "template %d: nodes/pods %s vs oracle %s" % (t, (nc[t], pc[t]), (onc, opc))
            assert np.array_equal(sched[t], osched), "template %d: per-group scheduled counts" % t
            assert np.array_equal(order[t], oorder), "template %d: processing order" % t
    # size-independent properties on the full result
    assert int(count.max()) <= enc.P and int(pc.max()) <= enc.P
    assert np.array_equal(sched.sum(axis=1), pc), "pod_count is the sum of the per-group scheduled counts"
    assert np.all((order >= -1) & (order < enc.E))
