#!/usr/bin/env python
"""bench.py -- throughput of the NRPT hot path on BASELINE.json config C2.

Workload (config.workload): Bayesian logistic regression, 100 000 synthetic rows x 32
features (seed 20240002), 64 chains per GPU, nPassesPerScan = 3, COPIES initialisation from
the prototype state w = 0 (steady-state exploration: every chain relaxes to its own annealed
posterior within a couple of scans; FORWARD initialisation from the N(0,10) prior leaves most
chains for thousands of scans on the reference's softened-support plateau, -beta*1e100 per
out-of-support row, which is a burn-in pathology of the reference algorithm, not the regime PT
spends its time in).
One "step" = one NRPT scan = local exploration of every chain (96 slice-sampler executions
per chain) + the DEO swap pass.  Metric = annealed log-density evaluations per second
(one evaluation = one sampler logDensity() call over all factors connected to the moved
variable, R/mcmc/RealSliceSampler.java:156-161; counted on the device).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU restatement of the reference engine

N > 1 is launched by torchrun, one rank per GPU: chain slots are sharded (64 chains per GPU,
one 64*N-chain ladder, weak scaling); the ranks exchange their mailbox handles once
(connect_peers), after which every chain's 32-byte record is stored into every rank's mailbox
by the exploration kernel itself -- no collective and no host code per scan -- and every rank
takes the same swap decisions.  Both `value` and `e2e` drive that sharded engine.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS, N_FEATURES, CHAINS_PER_GPU = 100_000, 32, 64
ALGO_BYTES_PER_EVAL = N_ROWS * (N_FEATURES * 8 + 1)      # one full pass over (x_i, y_i)
KERNEL_BYTES_PER_EVAL = N_ROWS * (8 + 8)                 # what the eta-cached kernel moves per evaluation: eta_i, X[i][k]
# fp64-pipe instructions (DFMA + DADD + DMUL [+ DSETP]) per row of one evaluation, counted in the SASS of the inner
# loop of each exploration kernel (profiles/sass_pool_inner_loop_r02.md, tools/sass_loops.py): the kernel's roofline is
# the fp64 pipe (one fp64 instruction per lane per 2 cycles per SM sub-partition), not HBM -- the rows come from L2.
FP64_INSTR_PER_ROW = {1: ("pool_kernel<LogisticPool>", 20), 0: ("explore_kernel<LogisticFam>", 24)}
NRPT_SCANS = 400                                          # scans of the adaptive NRPT run that measures round trips/s
METRIC = "annealed log-density evals/s (NRPT, logistic 100k x 32, 64 chains/GPU)"
UNIT = "evals/s"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned(a):
    """copy of `a` in page-locked host memory (numpy view of a pinned torch tensor)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle (CPU restatement of the reference engine)
# ---------------------------------------------------------------------------------------
def cpu_sample(spec, budget_s, n_chains=None):
    """Times the oracle on a bounded sample of the workload: full data, `n_chains` chains
    (one thread per chain), a fraction of a scan chosen to fit `budget_s`.  Returns
    (evals_per_second, cores, description, elapsed)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    if n_chains is None:
        n_chains = max(2, min(CHAINS_PER_GPU, cores))
    model = O.Model(spec)
    # calibration: one sampler execution per chain
    pt = O.PT(model, nChains=n_chains, seed=1, nPassesPerScan=1.0 / N_FEATURES, nThreads=cores, usePriorSamples=False)
    t0 = time.perf_counter(); pt.scan(); t_exec = time.perf_counter() - t0
    n_exec = int(max(1, min(3 * N_FEATURES, budget_s / max(t_exec, 1e-3))))
    pt2 = O.PT(model, nChains=n_chains, seed=2, nPassesPerScan=(n_exec + 0.5) / N_FEATURES, nThreads=cores,
               usePriorSamples=False)
    e0 = pt2.n_evals()
    t0 = time.perf_counter(); pt2.scan(); dt = time.perf_counter() - t0
    evals = pt2.n_evals() - e0
    desc = ("oracle port, full 100k x 32 data, COPIES init, %d chains x %d sampler executions each (of 96 per scan), "
            "%d threads" % (n_chains, n_exec, min(cores, n_chains)))
    return evals / dt, min(cores, n_chains), desc, dt


def workload(n_chains):
    """config.workload, the same string in both arms"""
    return ("C2 Bayesian logistic regression 100000 x 32 (seed 20240002), %d chains (%d per GPU), nPassesPerScan 3, "
            "COPIES init from w=0; step = one NRPT scan (explore + DEO swap)" % (n_chains, CHAINS_PER_GPU))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from blangsdk_b200 import datasets
    spec = datasets.logistic_regression(N_ROWS, N_FEATURES)
    total_budget = 150.0
    per_step = max(0.5, total_budget / max(1, args.steps + args.warmup))
    vals, cores, desc = [], 1, ""
    t_all = 0.0
    for s in range(args.warmup + args.steps):
        v, cores, desc, dt = cpu_sample(spec, per_step / 2.0)
        if s >= args.warmup:
            vals.append(v); t_all += dt
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(CHAINS_PER_GPU * max(1, args.gpus)),
                       "reference": "CPU restatement of the reference PT engine (oracle/, no JVM in the image), all host "
                                    "threads; each step is a bounded sample of one scan of this workload, see cpu_baseline.sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from blangsdk_b200 import datasets
    from blangsdk_b200.engine import Engine, microbench
    from blangsdk_b200.distributed import ShardedEngine, connect_peers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_chains = CHAINS_PER_GPU * world

    spec = datasets.logistic_regression(N_ROWS, N_FEATURES)
    eng = Engine(nChains=n_chains, random=1, device=local_rank, rank=rank, worldSize=world, ctasPerChain=args.ctas)
    eng.set_model(spec)
    if world > 1:
        connect_peers(eng)        # mailboxes over NVLink: from here on no host code runs between the scans' kernels
    eng.initialize("COPIES")
    sh = ShardedEngine(eng)
    stream = torch.cuda.Stream(device=dev)
    sh.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if world > 1:
        # the in-library exchange replaces the NCCL all-gather of ShardedEngine: explore publishes the chains' records
        # to every rank and completes this rank's table (gather_kernel), so the step is two C calls and no collective
        sh.exchange = lambda: None
        sh._gather = lambda: None
    with torch.cuda.stream(stream):
        sh.prepare()
        for _ in range(args.warmup):
            sh.scan()
        barrier()
        c0 = eng.counters()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        barrier()
        wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(s & 0xff)            # L2 flush between timed steps (outside the event pairs)
            ev[s][0].record(stream)
            sh.explore()
            ev[s][1].record(stream)
            sh.exchange()
            sh.swap()
            ev[s][2].record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        clk = clocks.stop() if rank == 0 else None
        c1 = eng.counters()
    sh.synchronize()
    ms_explore = sum(e[0].elapsed_time(e[1]) for e in ev)
    ms_total = sum(e[0].elapsed_time(e[2]) for e in ev)
    evals = c1["evals"] - c0["evals"]
    launches = c1["launches"] - c0["launches"]
    if world > 1:
        t = torch.tensor([ms_total, ms_explore], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_explore = float(t[0]), float(t[1])
        t = torch.tensor([float(evals), float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        evals, launches = int(t[0]), int(t[1])
    value = evals / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers: the SAME sharded engine shape as `value` ----
    e2e = None
    if not args.no_e2e:
        xs, _k1 = pinned(spec["x"]); ys, _k2 = pinned(spec["y"])
        spec_p = dict(spec, x=xs, y=ys)
        k_e2e = args.steps
        runs = []
        for _rep in range(5):                         # five complete repetitions, the median is reported:
            barrier()                                 # a driver hiccup in create/free (seen: +0.3 s) is not the path
            t0 = time.perf_counter()
            e2 = Engine(nChains=n_chains, random=2, device=local_rank, rank=rank, worldSize=world, ctasPerChain=args.ctas)
            t1 = time.perf_counter()
            e2.set_model(spec_p)                      # H2D of the data set (host pointers), on every rank
            t2 = time.perf_counter()
            if world > 1:
                connect_peers(e2)
            t3 = time.perf_counter()
            e2.initialize("COPIES")
            out = e2.run_scans(k_e2e, want=("energy", "swapIndicators", "samples", "logDensity", "nOutOfSupport"))
            cc = e2.counters()
            torch.cuda.synchronize(dev)
            t4 = time.perf_counter()
            runs.append((t4 - t0, cc["evals"], {"create": t1 - t0, "set_model": t2 - t1, "peer_connect": t3 - t2,
                                                "initialize_and_scans": t4 - t3}))
            barrier()                                 # nobody frees its mailbox while a peer may still write to it
            e2.close()
        runs.sort(key=lambda r: r[0])
        dt, n_evals, phases = runs[len(runs) // 2]
        d2h = sum(v.nbytes for v in out.values()) / k_e2e
        h2d = (xs.nbytes + ys.nbytes) / k_e2e
        e2e_value = n_evals / dt
        if world > 1:
            t = torch.tensor([dt, float(n_evals)], dtype=torch.float64, device=dev)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            e2e_value = float(tsum[1]) / float(tmax[0])
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "phases_s_rank0": phases,
               "what": "Engine(nChains=%d, worldSize=%d) create + set_model (data set H2D from pinned host memory, every rank) + "
                       "%sinitialize + run_scans(%d) with per-scan outputs copied D2H + counters; wall clock, max over ranks, "
                       "median of 5 complete repetitions" % (n_chains, world, "peer connect + " if world > 1 else "", k_e2e)}

    # ---- NRPT quality of the same workload: adaptive rounds, then round trips/s of the final round ----
    nrpt = None
    if not args.no_nrpt:
        e3 = Engine(nChains=n_chains, random=5, device=local_rank, rank=rank, worldSize=world, ctasPerChain=args.ctas)
        e3.set_model(spec)
        if world > 1:
            connect_peers(e3)
        e3.initialize("COPIES")
        t0 = time.perf_counter()
        res = e3.perform_inference(args.nrpt_scans, 0.5, want=())
        wall_nrpt = time.perf_counter() - t0
        last = res["rounds"][-1]
        nrpt = {"what": "PT.performInference(nScans=%d, adaptFraction=0.5), COPIES init, same workload; figures of the final "
                        "(non-adaptive) round: round trips = Paths.nRejuvenations (ptanalysis/Paths.xtend:26-50), rate = "
                        "count / that round's wall time (PT.java:451-461 reports the same count per scan)" % args.nrpt_scans,
                "rounds": [r["nScans"] for r in res["rounds"]], "final_round_scans": last["nScans"],
                "final_round_s": last["millis"] * 1e-3, "round_trips": int(last["roundTrips"][0]),
                "round_trips_per_s": float(last["roundTrips"][0]) / (last["millis"] * 1e-3),
                "Lambda": float(last["Lambda"][0]), "mean_swap_accept": float(last["meanAccept"][0]),
                "min_swap_accept": float(last["minAccept"][0]), "logZ_stepping_stone": float(last["logZ"][0]),
                "wall_s": wall_nrpt}
        st3 = e3.round_stats()
        dens3 = e3.log_density()
        nrpt["bottom_pair_accept"] = float(st3["swapAcceptMean"][0][-1])
        nrpt["n_out_of_support_at_beta0"] = int(dens3["noos"][0][-1])
        nrpt["note"] = ("round trips are 0 on this workload under the reference's own semantics: a draw from the N(0, 10) prior "
                        "gives margins x.w with standard deviation ~18, so thousands of the 100 000 naive Bernoulli factors "
                        "log(1 - p) are -inf (p rounds to 1.0 beyond 36.7); with annealSupport such a state has density "
                        "-beta * 1e100 * nOutOfSupport (ExponentiatedFactor.java:82-87), the swap between the beta = 0 chain "
                        "and its neighbour is never accepted (bottom_pair_accept), and no particle that visited beta = 0 can "
                        "reach beta = 1.  nrpt_other holds the same measurement on workloads whose prior draws are in support.")
        barrier()
        e3.close()
        # the same shape with the slab prior of the reference's own fixture (F/SpikedGLM.bl:22, Normal(0, 1)): prior draws
        # are in support, so particles do travel between beta = 0 and beta = 1 and the round-trip rate is not zero
        spec1 = datasets.logistic_regression(N_ROWS, N_FEATURES, prior_var=1.0)
        e4 = Engine(nChains=n_chains, random=6, device=local_rank, rank=rank, worldSize=world, ctasPerChain=args.ctas)
        e4.set_model(spec1)
        if world > 1:
            connect_peers(e4)
        e4.initialize("COPIES")
        t0 = time.perf_counter()
        res4 = e4.perform_inference(args.nrpt_scans_unit_prior, 0.5, want=())
        wall4 = time.perf_counter() - t0
        l4 = res4["rounds"][-1]
        nrpt["unit_prior"] = {
            "workload": "C2 shape (100000 x 32, %d chains) with w_j ~ Normal(0, 1), the slab of F/SpikedGLM.bl:22; "
                        "PT.performInference(nScans=%d, adaptFraction=0.5), COPIES init" % (n_chains, args.nrpt_scans_unit_prior),
            "final_round_scans": l4["nScans"], "final_round_s": l4["millis"] * 1e-3,
            "round_trips": int(l4["roundTrips"][0]),
            "round_trips_per_s": float(l4["roundTrips"][0]) / (l4["millis"] * 1e-3),
            "Lambda": float(l4["Lambda"][0]), "mean_swap_accept": float(l4["meanAccept"][0]),
            "min_swap_accept": float(l4["minAccept"][0]), "logZ_stepping_stone": float(l4["logZ"][0]), "wall_s": wall4,
            "min_scans_per_round_trip": 2 * (n_chains - 1),
            "note": "a particle moves one ladder position per scan at best, so one round trip takes at least 2 (nChains - 1) "
                    "scans: with the 64 N-chain ladder of N > 1 GPUs the final round of this run is shorter than one round trip "
                    "beyond N = 4"}
        barrier()
        e4.close()
        if rank == 0:        # BASELINE configs 0 and 2 (C1 at its named size, a down-scaled twin of C3) on this rank's GPU
            nrpt["nrpt_other"] = []
            for name, spec_o, chains_o, scans_o, init_o in (
                    ("C1 Normal 1000 obs, 8 chains, 1000 scans, SCM init", datasets.normal_meanvar(), 8, 1000, "SCM"),
                    ("C3 twin: mixture K=4, 20000 obs, 32 chains, 2000 scans, SCM init", datasets.gaussian_mixture(20000), 32, 2000, "SCM")):
                eo = Engine(nChains=chains_o, random=7, device=local_rank)
                eo.set_model(spec_o)
                eo.initialize(init_o)
                t0 = time.perf_counter()
                ro = eo.perform_inference(scans_o, 0.5, want=())
                wo = time.perf_counter() - t0
                lo = ro["rounds"][-1]
                nrpt["nrpt_other"].append({"workload": name, "final_round_scans": lo["nScans"], "final_round_s": lo["millis"] * 1e-3,
                                           "round_trips": int(lo["roundTrips"][0]),
                                           "round_trips_per_s": float(lo["roundTrips"][0]) / (lo["millis"] * 1e-3),
                                           "Lambda": float(lo["Lambda"][0]), "mean_swap_accept": float(lo["meanAccept"][0]),
                                           "min_swap_accept": float(lo["minAccept"][0]), "logZ_stepping_stone": float(lo["logZ"][0]),
                                           "wall_s": wo, "n_gpus": 1})
                eo.close()

    if rank == 0:
        peak, peak_src = read_peaks()
        evals_per_launch = evals / max(1, args.steps * world)
        explore_s = ms_explore * 1e-3 / args.steps
        try:
            fp64_peak = microbench(0, local_rank)          # measured DFMA rate of this GPU, TFLOP/s (2 flop per FMA)
        except Exception:
            fp64_peak = None
        kernel_name, per_row = FP64_INSTR_PER_ROW[eng.kernel_variant]
        rows_per_s = evals_per_launch * N_ROWS / explore_s
        achieved_tf = rows_per_s * per_row * 2 / 1e12        # every fp64-pipe instruction counted as one FMA issue slot
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        algo_gbps = evals_per_launch * ALGO_BYTES_PER_EVAL / explore_s / 1e9
        roofline = {"bound": "fp64", "kernel": kernel_name, "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (achieved_tf / fp64_peak) if fp64_peak else None,
                    "peak_source": "fp64 FMA rate measured on this GPU in this run (blangpt_microbench(0)); "
                                   "MEASURED_PEAKS.json has no fp64 figure",
                    "how": "rows/s x %d fp64-pipe instructions per row (DFMA+DADD+DMUL in the inner-loop SASS) x 2, over the "
                           "explore kernel's CUDA-event time; ncu sm__pipe_fp64_cycles_active of the same kernel: "
                           "profiles/pool_logistic_r02.md.  The count is of the work the kernel does, not of the work it "
                           "avoids: the last build of round 2 took 2 of 22 fp64 instructions out of the row term (4.01 -> 4.11 M "
                           "evals/s), which lowers this fraction (0.50 -> 0.47) while the throughput rises" % per_row,
                    "fp64_instr_per_row": per_row, "rows_per_s": rows_per_s, "evals_per_launch": evals_per_launch,
                    "ms_per_launch": explore_s * 1e3, "traffic": traffic, "traffic_source": traffic_src,
                    "hbm_side": {"note": "SURVEY 8(d) counts one full pass over X,y (25.7 MB) per evaluation; the engine keeps "
                                         "eta = X.w per chain, moves 16 B/row per evaluation from L2 and is NOT HBM-bound: the "
                                         "8(d) figure is far above 1 x HBM peak and is reported here only as the survey asks",
                                 "algorithmic_bytes_per_eval": ALGO_BYTES_PER_EVAL, "algorithmic_GBps": algo_gbps,
                                 "hbm_peak_GBps": peak, "hbm_peak_source": peak_src,
                                 "algorithmic_x_of_hbm_peak": algo_gbps / peak,
                                 "kernel_bytes_per_eval": KERNEL_BYTES_PER_EVAL,
                                 "kernel_L2_GBps": evals_per_launch * KERNEL_BYTES_PER_EVAL / explore_s / 1e9}}
        cpu_v, cpu_cores, cpu_desc, _ = cpu_sample(spec, 8.0) if not args.no_cpu else (None, 0, "skipped", 0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload(n_chains),
                           "l2": "flushed between timed steps (256 MiB write)", "ctas_per_chain": eng.ctas_per_chain,
                           "parallelism": "chain-sharded x%d" % world},
                "scans_per_s": 1e3 * args.steps / ms_total, "evals": evals, "wall_s_timed_region": wall,
                "kernel_variant": "pool" if eng.kernel_variant == 1 else "cluster",
                "round_trips_per_s": nrpt["unit_prior"]["round_trips_per_s"] if nrpt else None,
                "round_trips_workload": nrpt["unit_prior"]["workload"] if nrpt else None,
                "round_trips_per_s_baseline_prior": nrpt["round_trips_per_s"] if nrpt else None, "nrpt": nrpt,
                "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cpu_cores, "kind": "port", "sample": cpu_desc}}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    global CHAINS_PER_GPU, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--ctas", type=int, default=0, help="CTAs per chain (0 = auto)")
    ap.add_argument("--chains-per-gpu", type=int, default=CHAINS_PER_GPU,
                    help="default 64 = BASELINE config 1; 128 = the north-star shape (1024 chains on 8 GPUs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nrpt", action="store_true", help="skip the adaptive NRPT run that measures round trips/s")
    ap.add_argument("--nrpt-scans", type=int, default=NRPT_SCANS)
    ap.add_argument("--nrpt-scans-unit-prior", type=int, default=1600)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.chains_per_gpu != CHAINS_PER_GPU:
        CHAINS_PER_GPU = args.chains_per_gpu
        METRIC = METRIC.replace("64 chains/GPU", "%d chains/GPU" % CHAINS_PER_GPU)
    args.warmup = max(3, args.warmup) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
