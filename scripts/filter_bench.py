#!/usr/bin/env python
"""filter-out-schedulable (HintingSimulator.TrySchedulePods) on a synthetic config: GPU pass through the C ABI vs the
CPU oracle on a bounded prefix of the same pod order.

    python scripts/filter_bench.py --config 3 [--pods N --cluster-nodes N] [--cpu-pods 2000]
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
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--pods", type=int, default=None)
    ap.add_argument("--cluster-nodes", type=int, default=None)
    ap.add_argument("--cpu-pods", type=int, default=2000, help="prefix of the order timed on the CPU oracle (0 = skip)")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from kubernetes_autoscaler_b200 import synth
    from kubernetes_autoscaler_b200.engine import Engine
    enc = synth.generate(args.config, pods=args.pods, templates=8, cluster_nodes=args.cluster_nodes)
    order = np.arange(enc.P, dtype=np.int32)
    eng = Engine()
    rows = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        eng.load(enc)
        t1 = time.perf_counter()
        assigned, li, ov = eng.filter_schedulable(order)
        t2 = time.perf_counter()
        rows.append({"load_ms": 1e3 * (t1 - t0), "filter_wall_ms": 1e3 * (t2 - t1), "filter_dev_ms": eng.stats().estimate_ms})
    best = min(rows, key=lambda r: r["filter_wall_ms"])
    out = {"config": args.config, "pods": enc.P, "cluster_nodes": int(enc.arrays["num_cluster_nodes"]) if "num_cluster_nodes" in enc.arrays else None,
           "placed": int((assigned >= 0).sum()), "last_index": li, "best": best}
    if args.cpu_pods:
        from oracle import pyoracle
        k = min(args.cpu_pods, enc.P)
        t0 = time.perf_counter()
        want, wli, _ = pyoracle.filter_schedulable(enc, order[:k])
        dt = time.perf_counter() - t0
        got, gli, _ = eng.filter_schedulable(order[:k])
        out["cpu_oracle"] = {"pods": k, "seconds": dt, "parity": bool(np.array_equal(got, want) and gli == wli),
                             "extrapolated_seconds_all_pods": dt * enc.P / k, "cores": 1}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
