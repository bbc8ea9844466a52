#!/usr/bin/env python
"""Where does the dense pass's step time go?  Event-to-event time of cae_feasibility on C2 with a cold L2
(512 MiB memset before every step, as bench.py does), a warm L2, and of an empty torch kernel for the
launch + event floor of this box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from kubernetes_autoscaler_b200 import synth
    from kubernetes_autoscaler_b200.engine import Engine
    enc = synth.generate(2)
    eng = Engine()
    eng.load(enc)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, cold in (("cold", True), ("warm", False)):
        ms = []
        for i in range(25):
            if cold:
                flush.zero_()
            torch.cuda.synchronize()
            eng.lib.cae_feasibility(eng.h, None, None, None)
            if i >= 5:
                ms.append(eng.stats().feasibility_ms)
        out[name + "_us"] = 1e3 * float(np.mean(ms))
        out[name + "_min_us"] = 1e3 * float(np.min(ms))
    x = torch.zeros(1, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for i in range(25):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record()
        x.add_(1)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ms.append(e0.elapsed_time(e1))
    out["tiny_kernel_event_to_event_us"] = 1e3 * float(np.mean(ms))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
