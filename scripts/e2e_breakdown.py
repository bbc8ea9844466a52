#!/usr/bin/env python
"""Where the end-to-end time of the dense pass goes (C2, through the ABI with host buffers)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from kubernetes_autoscaler_b200 import synth
    from kubernetes_autoscaler_b200.engine import Engine
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    enc = synth.generate(cfg)
    eng = Engine()
    rows = []
    for i in range(12):
        t0 = time.perf_counter()
        eng.load(enc)
        t1 = time.perf_counter()
        st = eng.stats()
        h2d = st.h2d_ms
        eng.feasibility()
        t2 = time.perf_counter()
        st = eng.stats()
        if i >= 2:
            rows.append([1e3 * (t1 - t0), h2d, 1e3 * (t2 - t1), st.feasibility_ms, st.d2h_ms])
    a = np.array(rows)
    m = a.mean(axis=0)
    print(json.dumps({"config": cfg, "load_wall_ms": m[0], "load_h2d_ms": m[1], "feasibility_wall_ms": m[2],
                      "kernel_ms": m[3], "d2h_ms": m[4], "h2d_bytes": int(st.h2d_bytes), "d2h_bytes": int(st.d2h_bytes)}))


if __name__ == "__main__":
    main()
