"""Fused multi-GPU histogram exchange (cae_peer_attach): two ranks, one process per GPU, pods sharded;
fit_count of every rank must equal the single-rank / oracle histogram.  Needs >= 2 GPUs (skipped otherwise)."""
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from kubernetes_autoscaler_b200 import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, q_in, q_out, result):
    import __graft_entry__  # noqa: F401  (sys.path)
    from kubernetes_autoscaler_b200.engine import Engine, unpack_bits
    enc = synth.generate(2, pods=10_000, templates=200)
    eng = Engine(device=rank, rank=rank, world_size=world)
    q_out.put((rank, eng.peer_handle()))
    handles = q_in.get()
    eng.peer_attach(handles)
    eng.load(enc)
    out = []
    for _ in range(5):   # several steps: slots alternate and are cleared between uses
        bits, _, count = eng.feasibility()
        out.append(count.copy())
    pb, pe = eng.pod_shard(enc.P)
    result.put((rank, out, unpack_bits(bits, pe - pb).sum(axis=1)))
    q_in.get()  # keep the exchange buffer alive until every rank is done
    eng.close()


def test_two_rank_fused_exchange(oracle):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q_ins = [ctx.Queue() for _ in range(world)]
    q_out, result = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, q_ins[r], q_out, result)) for r in range(world)]
    for p in procs:
        p.start()
    handles = dict(q_out.get(timeout=120) for _ in range(world))
    for r in range(world):
        q_ins[r].put([handles[i] for i in range(world)])
    res = [result.get(timeout=300) for _ in range(world)]
    for r in range(world):
        q_ins[r].put("done")
    for p in procs:
        p.join(timeout=60)
    enc = synth.generate(2, pods=10_000, templates=200)
    want, _ = oracle.feasibility_dense(enc)
    full = (want == 0).sum(axis=1)
    local_sum = np.zeros_like(full)
    for rank, counts, local in res:
        for c in counts:
            assert np.array_equal(c, full), "rank %d: all-reduced histogram differs from the oracle" % rank
        local_sum += local
    assert np.array_equal(local_sum, full)
