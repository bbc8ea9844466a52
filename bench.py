#!/usr/bin/env python
"""bench.py — pod x node predicate evaluations/sec of the dense feasibility pass (BASELINE.json
metric 1) on config C2: 100 000 pods x 1 000 templates, resources + taints/tolerations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2]

One "step" = one dense pass of the scale-up predicate path over the whole pending-pod batch:
every pod (not only group exemplars) against every template through the Filter chain.
  value        : whole-job evals/s with the snapshot already resident in HBM (device time, CUDA events)
  e2e          : the same through the C-ABI with HOST buffers: cae_load (intern + H2D + class
                 matrices) + cae_feasibility (kernel + D2H of the bit matrix and counts) per step
  roofline     : feasibility_kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak
  cpu_baseline : the CPU oracle (port of the Go reference) on the box's host cores, same workload
N > 1 (torchrun): weak scaling — every rank owns 100 000 pods of an N x 100 000-pod snapshot, the
per-template fit-count histogram int32[T] is all-reduced once per step over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


_CLOCK_QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"


def _clock_sampler_start(devs):
    """ONE looping nvidia-smi for the whole timed region (the profiling recipe's clocks line), started by rank 0
    only: spawning nvidia-smi per sample from every rank stalls kernel launches on a multi-GPU box for milliseconds."""
    try:
        return subprocess.Popen(["nvidia-smi", "-i", ",".join(str(d) for d in devs), "--query-gpu=" + _CLOCK_QUERY,
                                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def _clock_sampler_stop(proc):
    if proc is None:
        return []
    try:
        proc.terminate()
        out, _ = proc.communicate(timeout=5)
    except Exception:
        proc.kill()
        return []
    return [[x.strip() for x in line.split(",")] for line in out.strip().splitlines() if line.count(",") >= 5]


def _clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max(int(s[1]) for s in samples if s[1].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def _peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


_W = {}


def _worker_init(config, pods, templates):
    from kubernetes_autoscaler_b200 import synth
    from oracle import pyoracle
    _W["enc"] = synth.generate(config, pods=pods, templates=templates)
    _W["oracle"] = pyoracle
    pyoracle.lib()


def _worker_run(job):
    p_range, t_range = job
    t0 = time.perf_counter()
    ev = _W["oracle"].feasibility_dense(_W["enc"], p_range=p_range, t_range=t_range)[1]
    return ev, time.perf_counter() - t0


def _worker_decide(job):
    tb, te, cap = job
    enc = _W["enc"]
    caps = np.full(enc.T, cap, np.int32)
    t0 = time.perf_counter()
    _W["oracle"].estimate_all(enc, caps, t_range=(tb, te))
    return time.perf_counter() - t0


def _worker_pid(_):
    time.sleep(0.02)
    return os.getpid()


def _wait_workers(pool, procs):
    """Block until every pool worker has finished its initializer (generated its snapshot copy)."""
    for _ in range(200):
        if len(set(pool.map(_worker_pid, range(4 * procs), chunksize=1))) >= procs:
            return


def _cpu_dense(pool, procs, p_range, t_range):
    """Oracle dense feasibility on `procs` host processes (templates split across them).
    Returns (evals, wall seconds)."""
    tb, te = t_range
    cuts = [tb + (te - tb) * i // procs for i in range(procs + 1)]
    jobs = [(p_range, (cuts[i], cuts[i + 1])) for i in range(procs) if cuts[i + 1] > cuts[i]]
    t0 = time.perf_counter()
    res = pool.map(_worker_run, jobs)
    return int(sum(r[0] for r in res)), time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--pods", type=int, default=None)
    ap.add_argument("--templates", type=int, default=None)
    ap.add_argument("--decision", action="store_true", help="reference arm: time full scale-up decisions (oracle) instead")
    ap.add_argument("--no-decision", action="store_true", help="engine arm: skip the secondary decision-latency figure")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: how the int32[T] fit histogram is reduced: fused P2P atomics in the kernel's last block, or NCCL")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)

    from kubernetes_autoscaler_b200 import synth
    cfg = synth.CONFIGS[args.config]
    P1 = args.pods or cfg.pods
    T = args.templates or cfg.templates
    workload = "%s; %d pods/GPU x %d templates, splitmix64 seed 0xCA5CA1E0+%d" % (cfg.name, P1, T, cfg.index)
    metric = "pod x node predicate evals/sec"

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        import multiprocessing as mp
        from oracle import pyoracle
        pyoracle.build()
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        P = P1
        # bounded sample per step: all pods x a template slice (~0.2 s of work per core)
        t_slice = min(T, 8 * cores)
        if args.decision:
            # metric 2: SchedulablePodGroups + Estimate per template, one template per job
            with mp.get_context("fork").Pool(cores, initializer=_worker_init, initargs=(args.config, P1, T)) as pool:
                _wait_workers(pool, cores)
                k = min(T, 2 * cores)
                t0 = time.perf_counter()
                per = pool.map(_worker_decide, [(t, t + 1, 1000) for t in range(k)], chunksize=1)
                wall = time.perf_counter() - t0
            print(json.dumps({"impl": "reference", "decision": True, "templates_timed": k, "wall_s": wall,
                              "cpu_seconds_per_template": float(np.mean(per)), "cores": cores,
                              "extrapolated_s_all_templates": wall * T / k, "kind": "port"}))
            return
        with mp.get_context("fork").Pool(cores, initializer=_worker_init, initargs=(args.config, P1, T)) as pool:
            _wait_workers(pool, cores)
            for _ in range(2):
                _cpu_dense(pool, cores, (0, P), (0, min(T, t_slice)))
            evals = 0
            secs = 0.0
            for s in range(args.steps):
                tb = (s * t_slice) % max(T - t_slice + 1, 1)
                ev, dt = _cpu_dense(pool, cores, (0, P), (tb, tb + t_slice))
                evals += ev
                secs += dt

        class _E:
            pass
        enc = _E()
        enc.P = P
        v = evals / secs
        sample = "%d pods x %d templates per step (template slice of the full workload), %d steps" % (enc.P, t_slice, args.steps)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload, "note": "CPU oracle = C++ port of the Go reference (no Go toolchain in the image)"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ engine arm
    import torch
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    from kubernetes_autoscaler_b200.engine import Engine

    enc = synth.generate(args.config, pods=P1 * world, templates=T)   # weak scaling: P1 pods per rank
    eng = Engine(device=local_rank, rank=rank, world_size=world, want_reasons=False)
    eng.load(enc)
    pb, pe = eng.pod_shard(enc.P)
    Pl = pe - pb
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    count_t = None
    fused = False
    if world > 1 and args.collective == "peer":
        try:
            handles = [None] * world
            dist.all_gather_object(handles, eng.peer_handle())
            eng.peer_attach(handles)
            dist.barrier()
            fused = True
        except Exception as ex:   # no P2P between these devices: fall back to the NCCL all-reduce
            if rank == 0:
                print("peer exchange unavailable (%r); using NCCL" % (ex,), file=sys.stderr)
    if world > 1 and not fused:
        ptr, nbytes = eng.device_buffer(0)

        class _Wrap:
            __cuda_array_interface__ = {"shape": (T,), "typestr": "<i4", "data": (ptr, False), "version": 3}
        count_t = torch.as_tensor(_Wrap(), device="cuda")

    ar0 = torch.cuda.Event(enable_timing=True)
    ar1 = torch.cuda.Event(enable_timing=True)
    # The pass is a single ~10 us kernel: timed right after a host synchronize, the event window would mostly hold
    # the HOST's launch latency (the GPU idles between the first event and the kernel's arrival: ~12 us for an
    # empty kernel on this box, scripts/k1_floor.py).  So the L2 flush and a short spin kernel are queued on the
    # engine's own stream first; event, kernel and event are then enqueued while the GPU is still busy and the
    # window measures device time only.  wall_ms_per_step keeps the host view.
    estream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))

    sync_t = torch.zeros(1, device="cuda")

    def flush_l2():
        with torch.cuda.stream(estream):
            flush.zero_()                                      # > L2: evicts everything the previous step left
            torch.cuda._sleep(300_000)                         # ~150 us spin: covers the host's enqueue of the step
            if dist is not None:
                dist.all_reduce(sync_t)                        # device-side barrier on the engine's stream: the ranks' timed
                                                               # windows open together (a host barrier cannot align queued work)

    def step_resident():
        eng.lib.cae_feasibility(eng.h, None, None, None)       # kernel only; results stay in HBM
        if count_t is not None:
            ar0.record()
            dist.all_reduce(count_t)                           # int32[T] histogram over NVLink
            ar1.record()

    for _ in range(args.warmup):
        flush_l2()
        step_resident()
    torch.cuda.synchronize()

    sampler = _clock_sampler_start([local_rank]) if rank == 0 else None   # rank 0's GPU only: NVML queries delay launches
    if sampler is not None:
        time.sleep(0.5)            # nvidia-smi initialises NVML on every GPU of the box: keep that out of the timed steps
    launches0 = eng.stats().kernel_launches
    dev_ms, wall_ms, ar_ms = [], [], []
    for _ in range(args.steps):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        flush_l2()                                             # L2 flush between timed iterations (untimed, same stream)
        step_resident()
        torch.cuda.synchronize()
        dev_ms.append(eng.stats().feasibility_ms)
        if count_t is not None:
            ar_ms.append(ar0.elapsed_time(ar1))
    launches = eng.stats().kernel_launches - launches0
    for _ in range(10):                                        # host view of a step: launch + device + synchronize
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_resident()
        torch.cuda.synchronize()
        wall_ms.append(1e3 * (time.perf_counter() - t0))
    kern_ms = float(np.mean(dev_ms))
    # device time of a step: the pass (CUDA events on the engine's stream) + for N>1 the NCCL all-reduce of
    # the histogram (CUDA events on torch's stream), max over ranks
    allreduce_ms = float(np.mean(ar_ms)) if ar_ms else 0.0
    step_ms = kern_ms + allreduce_ms
    if dist is not None:
        tt = torch.tensor([step_ms, kern_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms, kern_ms = float(tt[0]), float(tt[1])
    value = (P1 * world) * T / (step_ms * 1e-3)

    # ---- e2e through the C ABI with host buffers: load (H2D) + pass + D2H of bits/counts ----------
    e2e_ms = []
    h2d = d2h = 0
    for i in range(max(3, min(args.steps, 10)) + 1):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        eng.load(enc)
        bits, _, cnt = eng.feasibility()
        if count_t is not None:
            dist.all_reduce(count_t)
            torch.cuda.synchronize()
        dt = 1e3 * (time.perf_counter() - t0)
        if i > 0:
            e2e_ms.append(dt)
        st = eng.stats()
        h2d, d2h = st.h2d_bytes, st.d2h_bytes
    e2e_step = float(np.mean(e2e_ms))
    if dist is not None:
        tt = torch.tensor([e2e_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_step = float(tt[0])
    samples = _clock_sampler_stop(sampler)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (feasibility_lut_kernel) -----------------------------------------
    # algorithmic bytes of the dense kernel: per pod W packed-rank words + 2 class ids,
    # per template W words, the bit matrix, the fit histogram
    req = enc.arrays["ps_req"][np.unique(enc.arrays["pend_spec"])]
    bits = 0
    for a in range(req.shape[1]):
        dv = len(np.unique(req[:, a][req[:, a] > 0]))
        if dv:
            bits += int(dv).bit_length() + 1
    Wd = max(1, (bits + 31) // 32)
    alg_bytes = Pl * (4 * Wd + 8) + T * 4 * Wd + Pl * T // 8 + 4 * T
    peak, peak_src = _peak_hbm()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one `ncu --set full` capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json")))
        if args.config == 2 and P1 == 100_000 and T == 1000:
            traffic = int(tj["dram__bytes_read.sum"]) + int(tj["dram__bytes_write.sum"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "feasibility_lut_kernel", "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "note": "shuffle / shared-memory issue bound: ~1 bit of compulsory HBM traffic per evaluation (DESIGN.md §K1)"}

    # ---- CPU baseline: the oracle on this box's cores, bounded sample of the same workload ---------------
    # (a fresh process: the oracle's worker pool must fork before any CUDA context exists)
    cpu = {"value": None, "unit": "evals/s", "cores": 0, "kind": "port", "sample": "failed"}
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4",
                              "--config", str(args.config), "--pods", str(P1), "--templates", str(T)],
                             capture_output=True, text=True, timeout=600, env={k: v for k, v in os.environ.items()
                                                                              if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as ex:  # the bench line must still be printed
        cpu["sample"] = "failed: %r" % (ex,)

    # ---- metric 2: scale-up decision latency @ 100k pods x 5k templates (C3), one GPU ---------------------
    decision = None
    if world == 1 and not args.no_decision:
        try:
            enc3 = synth.generate(3)
            caps = np.full(enc3.T, 1000, np.int32)
            ms = []
            for i in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                eng.load(enc3)                                                 # intern + H2D + class / counter tables
                nc, pc, _, _ = eng.estimate_all(caps, want_sched=False, copy=False)  # exemplar feasibility, order, pack
                mask, _ = eng.expander_best([0, 1, 2], nc, pc)                 # least-waste, most-pods, least-nodes
                if i:
                    ms.append(1e3 * (time.perf_counter() - t0))
            st = eng.stats()
            decision = {"workload": synth.CONFIGS[3].name + ", node cap 1000 per template", "ms": float(np.median(ms)),
                        "estimate_device_ms": st.estimate_ms, "nodes_total": int(nc.sum()), "pods_scheduled_total": int(pc.sum()),
                        "options_surviving_chain": int(mask.sum()), "cpu_baseline": None}
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--decision", "--config", "3"],
                                 capture_output=True, text=True, timeout=900,
                                 env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            decision["cpu_baseline"] = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as ex:
            decision = {"error": repr(ex)}

    print(json.dumps({
        "metric": metric, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload, "global_pods": P1 * world, "templates": T, "parallelism": "pods sharded x%d" % world,
                   "l2": "flushed between timed iterations (512 MiB memset on the engine's stream)",
                   "timing": "CUDA events on the engine's stream, queued behind the flush + a spin kernel (no host launch latency in the window)"},
        "kernel_ms": kern_ms, "allreduce_ms": allreduce_ms,
        "step_ms_rank0": {"min": float(np.min(dev_ms)), "median": float(np.median(dev_ms)), "max": float(np.max(dev_ms))},
        "collective": ("none" if world == 1 else ("fused P2P atomics over NVLink (peer memory)" if fused else "NCCL all_reduce int32[T]")), "wall_ms_per_step": float(np.mean(wall_ms)), "clocks": _clocks_summary(samples),
        "e2e": {"value": (P1 * world) * T / (e2e_step * 1e-3), "unit": "evals/s", "ms_per_step": e2e_step,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "decision_latency": decision}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
