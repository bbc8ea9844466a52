#!/usr/bin/env python
"""One full scale-up decision (load -> exemplar feasibility -> order -> pack -> expander) on a synthetic
config, with device timings per phase.  Used for profiling (ncu) and the decision-latency figure.

    python scripts/tick.py --config 2 [--pods N --templates N] [--cap 1000] [--reps 5] [--check]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--pods", type=int, default=None)
    ap.add_argument("--templates", type=int, default=None)
    ap.add_argument("--cluster-nodes", type=int, default=None)
    ap.add_argument("--cap", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", type=int, default=0, help="compare the first N templates with the CPU oracle")
    args = ap.parse_args()
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:   # torchrun: templates sharded over the ranks; all-reduce of int32[2T] (node_count | pod_count) + float64[T] (waste)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as ge
    ge.build()
    from kubernetes_autoscaler_b200 import synth
    from kubernetes_autoscaler_b200.engine import Engine
    t0 = time.perf_counter()
    enc = synth.generate(args.config, pods=args.pods, templates=args.templates, cluster_nodes=args.cluster_nodes)
    gen_s = time.perf_counter() - t0
    eng = Engine(device=local_rank, rank=rank, world_size=world)
    caps = np.full(enc.T, args.cap, np.int32)
    counts_t = None
    rows = []
    for rep in range(args.reps):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        eng.load(enc)
        t1 = time.perf_counter()
        nc, pc, sched, order = eng.estimate_all(caps, want_sched=dist is None or args.check > 0, copy=False)
        if dist is not None:
            if counts_t is None:
                ptr, nbytes = eng.device_buffer(1)

                class _Wrap:
                    __cuda_array_interface__ = {"shape": (2 * enc.T,), "typestr": "<i4", "data": (ptr, False), "version": 3}
                counts_t = torch.as_tensor(_Wrap(), device="cuda")
            waste_t = torch.from_numpy(eng.waste_scores()).cuda()   # own rows, 0.0 elsewhere (before the counts are summed)
            dist.all_reduce(counts_t)                                # int32[2T]: node_count | pod_count
            dist.all_reduce(waste_t)                                 # float64[T]: exactly one non-zero contribution per row
            torch.cuda.synchronize()
            both = counts_t.cpu().numpy()
            nc, pc = both[:enc.T].copy(), both[enc.T:].copy()
        t2 = time.perf_counter()
        if dist is not None:
            from kubernetes_autoscaler_b200.engine import expander_chain
            waste = waste_t.cpu().numpy()
            mask = expander_chain([0, 1, 2], nc, pc, waste)
        else:
            mask, waste = eng.expander_best([0, 1, 2], nc, pc)
        t3 = time.perf_counter()
        st = eng.stats()
        rows.append({"load_ms": 1e3 * (t1 - t0), "estimate_wall_ms": 1e3 * (t2 - t1), "estimate_dev_ms": st.estimate_ms,
                     "expander_wall_ms": 1e3 * (t3 - t2), "decision_ms": 1e3 * (t3 - t0)})
    best = min(rows, key=lambda r: r["decision_ms"])
    out = {"config": args.config, "pods": enc.P, "templates": enc.T, "groups": enc.E, "cap": args.cap, "gen_s": gen_s,
           "best": best, "median_decision_ms": float(np.median([r["decision_ms"] for r in rows])),
           "nodes_total": int(nc.sum()), "pods_scheduled_total": int(pc.sum()), "best_options": int(mask.sum())}
    if args.check:
        from oracle import pyoracle
        n = min(args.check, enc.T if dist is None else eng.template_shard(enc.T)[1])   # order rows exist on the owner only
        t0 = time.perf_counter()
        onc, opc, osched, oorder, ev = pyoracle.estimate_all(enc, caps, t_range=(0, n))
        out["oracle_s_for_%d_templates" % n] = time.perf_counter() - t0
        out["parity"] = bool(np.array_equal(nc[:n], onc) and np.array_equal(pc[:n], opc) and
                             np.array_equal(sched[:n], osched) and np.array_equal(order[:n], oorder))
    out["n_gpus"] = world
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
