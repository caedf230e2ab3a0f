"""Generates tests/golden/nrpt_golden.json from the CPU oracle (oracle/blang_oracle.c).

The reference ships no numeric golden vector for the NRPT path and no JVM exists in the
build container, so these vectors are produced by the oracle -- which is itself pinned on the reference's
known answers by tests/test_oracle_pins.py.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blangsdk_b200 import datasets as D  # noqa: E402
from oracle import oracle as O  # noqa: E402


def spec_to_json(spec):
    out = {"family": spec["family"], "hyper": spec["hyper"]}
    if "truth" in spec:
        out["truth"] = spec["truth"]
    for k in ("y", "x", "ntrials", "trials"):
        if spec.get(k) is not None:
            out[k] = np.asarray(spec[k]).tolist()
    return out


def main():
    specs = {
        "normal": D.normal_meanvar(n=40),
        "logistic": D.logistic_regression(60, 4),
        "mixture": D.gaussian_mixture(50),
        "betabinomial": D.beta_binomial(1, groups=8),
        "ising": D.ising(4),
    }
    # family 6 (appended after the first five so that their random states, and so their vectors, are unchanged)
    iid_state = {"poisson": lambda r: [r.gamma(2.0, 2.0)], "binomial": lambda r: [r.uniform(0.05, 0.95)],
                 "negbinomial": lambda r: [r.gamma(2.0, 1.5), r.uniform(0.05, 0.95)],
                 "studentt": lambda r: [r.gamma(2.0, 2.0), r.normal(1.0, 2.0), r.gamma(2.0, 1.0)],
                 "geometric": lambda r: [r.uniform(0.05, 0.95)], "laplace": lambda r: [r.normal(0.0, 2.0), r.gamma(2.0, 1.0)]}
    for dist in iid_state:
        specs["iid_" + dist] = D.iid_observations(dist, 45)
    # nine more laws; their runs start from the generating parameters ("start"), see tests/test_gpu_iid.py
    more = ["exponential", "gamma", "weibull", "gumbel", "logisticdist", "halfstudentt", "loglogistic", "gompertz", "yulesimon"]
    for dist in more:
        specs["iid_" + dist] = D.iid_observations(dist, 45)
    specs["hierarchical"] = D.hierarchical_rockets(6)
    specs["linreg"] = D.linear_regression(50, 3)
    specs["rockets_latent"] = D.rockets_latent(5)
    # appended last (earlier vectors unchanged): F/Doomsday.bl with the arguments of the reference's Getting Started transcript
    specs["iid_uniformmax"] = D.doomsday(1.2, 1.0)
    rng = np.random.default_rng(20240600)
    cases = []
    for name, spec in specs.items():
        m = O.Model(spec)
        dens = []
        for beta in (1.0, 0.37, 0.0):
            if name == "ising":
                re, it = None, rng.integers(0, 2, m.n_ints).tolist()
            elif name == "mixture":
                K = m.n_reals // 3
                re = np.concatenate([rng.normal(0, 3, K), rng.gamma(2, 1, K), rng.dirichlet(np.ones(K))]).tolist()
                it = None
            elif name == "logistic":
                re, it = rng.normal(0, 1.5, m.n_reals).tolist(), None
            elif name == "rockets_latent":
                G = m.n_ints
                re = rng.gamma(2.0, 1.0, 2).tolist() + rng.uniform(0.02, 0.98, G).tolist()
                it = (np.maximum(np.asarray(spec["y"]) - 1, 0) + rng.poisson(1.0, G)).tolist()
            elif name == "linreg":
                re, it = rng.normal(0, 2.0, m.n_reals - 1).tolist() + [float(rng.uniform(0.1, 5.0))], None
            elif name == "hierarchical":
                re, it = rng.gamma(2.0, 1.0, 2).tolist() + rng.uniform(0.02, 0.98, m.n_reals - 2).tolist(), None
            elif name.startswith("iid_") and name[4:] in iid_state:
                re, it = [float(v) for v in iid_state[name[4:]](rng)], None
            elif name.startswith("iid_"):
                re, it = [float(v * (1.0 + 0.2 * rng.normal())) for v in spec["truth"]], None
            else:
                re, it = rng.gamma(2.0, 1.0, m.n_reals).tolist(), None
            d = m.log_density(re, it, beta=beta)
            dens.append(dict(beta=beta, reals=re, ints=it, logdens=d["logdens"], loglik=d["loglik"],
                             noos=d["noos"], fixed=d["fixed"]))
        pt = O.PT(m, nChains=4, seed=2024,
                  sweepOrder=O.ORDER_CHECKERBOARD if name == "ising" else O.ORDER_REFERENCE)
        start = None
        if name == "rockets_latent":          # a feasible start (1 + L_g >= failures_g), see tests/test_gpu_rockets_latent.py
            G = m.n_ints
            start = []
            for p in range(4):
                re = (rng.gamma(2.0, 0.5, 2) + 0.2).tolist() + rng.uniform(0.1, 0.9, G).tolist()
                it = (np.maximum(np.asarray(spec["y"]) - 1, 0) + rng.poisson(1.0, G)).tolist()
                pt.set_state(p, re, it)
                start.append(re + [float(v) for v in it])      # the engine's layout: the integers after the reals
        elif name.startswith("iid_") and name[4:] not in iid_state:
            start = [[float(v * (1.0 + 0.1 * rng.normal())) for v in spec["truth"]] for _ in range(4)]
            for p, st in enumerate(start):
                pt.set_state(p, st)
        else:
            pt.initialize()
        res = pt.perform_inference(40, 0.5, want_samples=False)
        st = pt.round_stats()
        cases.append(dict(name=name, spec=spec_to_json(spec), densities=dens,
                          run=dict(nChains=4, seed=2024, nScans=40, adaptFraction=0.5, start=start,
                                   energies=res["energies"].tolist(), indicators=res["indicators"].tolist(),
                                   ladder=pt.get_ladder().tolist(), nEvals=pt.n_evals(),
                                   Lambda=st["Lambda"], logZ=st["log_z_stepping_stone"],
                                   roundTrips=st["round_trips"],
                                   swapAcceptMean=st["swap_accept_mean"].tolist())))
    with open(os.path.join(ROOT, "tests", "golden", "nrpt_golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", cases=cases), f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
