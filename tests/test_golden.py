"""Committed golden vectors (tests/golden/nrpt_golden.json, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU (-m gpu): the CUDA engine reproduces them through the C ABI."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-9


def _cases():
    with open(os.path.join(HERE, "golden", "nrpt_golden.json")) as f:
        return json.load(f)["cases"]


def _spec(c):
    s = dict(c["spec"])
    for k in ("y", "x", "ntrials", "trials"):
        if k in s:
            s[k] = np.array(s[k])
    if s["family"] == "linreg":
        s["y"] = s["y"].astype(np.float64)
    if s["family"] == "iid":
        s["hyper"] = dict(s["hyper"], priors=[tuple(p) for p in s["hyper"]["priors"]])
        s.setdefault("trials", None)
    return s


def _close(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.all((a == b) | (np.abs(a - b) <= TOL * np.maximum(1.0, np.abs(b))))


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_golden(oracle, case):
    spec = _spec(case)
    m = oracle.Model(spec)
    for d in case["densities"]:
        got = m.log_density(d["reals"], d["ints"], beta=d["beta"])
        assert got["noos"] == d["noos"] and _close(got["loglik"], d["loglik"]) and _close(got["logdens"], d["logdens"])
    r = case["run"]
    pt = oracle.PT(m, nChains=r["nChains"], seed=r["seed"],
                   sweepOrder=oracle.ORDER_CHECKERBOARD if case["name"] == "ising" else oracle.ORDER_REFERENCE)
    if r.get("start"):
        for p, st in enumerate(r["start"]):
            if case["name"] == "rockets_latent":
                pt.set_state(p, st[:m.n_reals], [int(v) for v in st[m.n_reals:]])
            else:
                pt.set_state(p, st)
    else:
        pt.initialize()
    res = pt.perform_inference(r["nScans"], r["adaptFraction"], want_samples=False)
    assert np.array_equal(res["indicators"], np.array(r["indicators"]))
    assert _close(res["energies"], r["energies"]) and _close(pt.get_ladder(), r["ladder"])
    assert pt.n_evals() == r["nEvals"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_cuda_reproduces_golden(case):
    from blangsdk_b200.engine import Engine, PT
    spec = _spec(case)
    eng = Engine(nChains=len(case["densities"]), random=1)
    eng.set_model(spec)
    eng.set_ladder(sorted([d["beta"] for d in case["densities"]], reverse=True))
    order = sorted(range(len(case["densities"])), key=lambda i: -case["densities"][i]["beta"])
    for p, i in enumerate(order):
        d = case["densities"][i]
        if case["name"] == "rockets_latent":      # the engine carries the latent counts after the reals
            eng.set_state(p, d["reals"] + [float(v) for v in d["ints"]], None)
        else:
            eng.set_state(p, d["reals"], d["ints"])
    got = eng.log_density()
    for p, i in enumerate(order):
        d = case["densities"][i]
        assert got["noos"][0, p] == d["noos"]
        assert _close(got["loglik"][0, p], d["loglik"]) and _close(got["fixed"][0, p], d["fixed"])
        assert _close(got["logdens"][0, p], d["logdens"])
    eng.close()
    r = case["run"]
    pt = PT(nChains=r["nChains"], nScans=r["nScans"], adaptFraction=r["adaptFraction"], random=r["seed"],
            initialization="COPIES" if r.get("start") else "FORWARD")
    pt.setSampledModel(spec)
    if r.get("start"):
        for p, st in enumerate(r["start"]):
            pt.engine.set_state(p, st, None)
        pt.engine.initialize("COPIES")
    res = pt.performInference()
    assert np.array_equal(res["swapIndicators"][:, 0, :], np.array(r["indicators"]))
    assert _close(res["energy"][:, 0, :], r["energies"]) and _close(res["ladder"][0], r["ladder"])
    assert res["counters"]["evals"] == r["nEvals"]
    st = res["finalStats"]
    assert int(st["roundTrips"][0]) == r["roundTrips"]
    assert _close(st["Lambda"][0], r["Lambda"]) and _close(st["logZSteppingStone"][0], r["logZ"])
    assert _close(st["swapAcceptMean"][0], r["swapAcceptMean"])
    pt.engine.close()
