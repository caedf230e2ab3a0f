"""Throughput of every BASELINE.json config shape on one GPU (C1..C5), a few scans each.
Writes one JSON object per config to stdout; used for profiles/configs_r01.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blangsdk_b200 import datasets as D  # noqa: E402
from blangsdk_b200.engine import Engine  # noqa: E402


def run(name, spec, nChains, nReplicas=1, init="FORWARD", warm=3, scans=10, algo_bytes=None, start=None):
    eng = Engine(nChains=nChains, nReplicas=nReplicas, random=1)
    eng.set_model(spec)
    if start is not None:                 # every chain from the same well-conditioned point
        for p in range(nChains):
            eng.set_state(p, start, None)
    eng.initialize(init)
    eng.run_scans(warm, want=())
    c0 = eng.counters()
    t0 = time.perf_counter()
    eng.run_scans(scans, want=())
    st_mid = eng.counters()
    dt = time.perf_counter() - t0
    ev = st_mid["evals"] - c0["evals"]
    out = dict(config=name, nChains=nChains, nReplicas=nReplicas, scans=scans, ms_per_scan=1e3 * dt / scans,
               evals_per_s=ev / dt, scans_per_s=scans / dt, evals_per_scan=ev / scans,
               ctas_per_chain=eng.ctas_per_chain, launches=st_mid["launches"] - c0["launches"])
    if algo_bytes:
        out["algorithmic_GBps"] = ev / dt * algo_bytes / 1e9
    # a short adaptive run for round trips / Lambda
    eng.reset_round()
    t0 = time.perf_counter()
    n_rt = max(scans * 4, 40)
    eng.run_scans(n_rt, want=())
    dt2 = time.perf_counter() - t0
    st = eng.round_stats()
    out.update(round_trips=int(st["roundTrips"].sum()), round_trips_per_s=float(st["roundTrips"].sum()) / dt2,
               Lambda=float(np.mean(st["Lambda"])))
    eng.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]
    if "C1" in which:
        run("C1 normal 1000 obs, 8 chains", D.normal_meanvar(), 8, scans=200, warm=20, algo_bytes=8000)
    if "C2" in which:
        run("C2 logistic 100k x 32, 64 chains", D.logistic_regression(), 64, init="COPIES", scans=10, algo_bytes=25_700_000)
    if "C2x1024" in which:   # north_star target regime: >= 1024 chains, eta cache (819 MB) streams from HBM
        run("C2 logistic 100k x 32, 1024 chains on one GPU", D.logistic_regression(), 1024, init="COPIES", scans=2, warm=2,
            algo_bytes=25_700_000)
    if "C2x128" in which:    # one GPU's share of 1024 chains sharded over 8 GPUs
        run("C2 logistic 100k x 32, 128 chains on one GPU", D.logistic_regression(), 128, init="COPIES", scans=5, warm=3,
            algo_bytes=25_700_000)
    for nc in (256, 512):    # where the pool kernel hands over to the cluster kernel (engine.cu, set_model)
        if "C2x%d" % nc in which:
            run("C2 logistic 100k x 32, %d chains on one GPU" % nc, D.logistic_regression(), nc, init="COPIES", scans=3, warm=2,
                algo_bytes=25_700_000)
    if "C3" in which:
        run("C3 mixture K=4 1M obs, 256 chains", D.gaussian_mixture(), 256, init="COPIES", scans=2, warm=1, algo_bytes=8_000_000)
    if "C3post" in which:    # every chain starts at the generating parameters: the colder chains stay near the posterior
        run("C3 mixture K=4 1M obs, 256 chains, started at the generating parameters", D.gaussian_mixture(), 256, init="COPIES",
            scans=3, warm=3, algo_bytes=8_000_000,
            start=[-6.0, -2.0, 2.0, 6.0, 1.0, 0.25, 0.25, 1.0, 0.25, 0.25, 0.25, 0.25])
    if "C4" in which:
        run("C4 ising 256x256, 128 chains (1 of 8 GPUs)", D.ising(256), 128, scans=20, warm=3, algo_bytes=None)
    if "C5" in which:
        run("C5 beta-binomial 4096 replicas x 8 chains", D.beta_binomial_batch(4096), 8, nReplicas=4096, scans=50, warm=5, algo_bytes=256)
    # the families added for SURVEY 8(f)-4 (not BASELINE configs): the same shapes as C2 / C3 / C5
    if "C6" in which:
        s = D.linear_regression(100_000, 32)
        run("C6 linreg 100k x 32, 64 chains", s, 64, init="COPIES", scans=10, algo_bytes=100_000 * (32 * 8 + 8), start=s["truth"])
    if "C7" in which:
        s = D.iid_observations("studentt", 1_000_000)
        run("C7 iid Student-t 1M obs, 64 chains", s, 64, init="COPIES", scans=5, warm=2, algo_bytes=8_000_000, start=s["truth"])
    if "C8" in which:
        run("C8 hierarchical rockets 12 groups, 4096 replicas x 8 chains", D.hierarchical_rockets(12), 8, nReplicas=4096,
            scans=50, warm=5, algo_bytes=12 * 8)
