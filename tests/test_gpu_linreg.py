"""-m gpu parity tests for family 8: F/LinRegression.bl with p covariates -- w_j ~ Normal(0, v0),
sigma ~ ContinuousUniform(0, smax), y_i ~ Normal(x_i . w, sigma * sigma)."""
import math

import numpy as np
import pytest

from gpu_util import assert_trajectory_parity, make_pair, rel, start_from

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_linreg_log_density_parity(oracle, datasets):
    from blangsdk_b200.engine import Engine
    spec = datasets.linear_regression(1501, 4)
    n = 6
    eng = Engine(nChains=n, random=1)
    eng.set_model(spec)
    m = oracle.Model(spec)
    assert (eng.n_reals, eng.n_samplers) == (m.n_reals, m.n_samplers) == (5, 5)
    rng = np.random.default_rng(3)
    states = [list(rng.normal(0, 2, 4)) + [s] for s in (0.7, 2.5, -1.1, 0.0, 11.0, 0.01)]   # sigma < 0: variance > 0 but prior -inf
    for p in range(n):
        eng.set_state(p, states[p], None)
    got = eng.log_density()
    ladder = eng.get_ladder()[0]
    for p in range(n):
        ref = m.log_density(states[p], None, beta=ladder[p])
        assert got["noos"][0, p] == ref["noos"]
        assert rel(got["loglik"][0, p], ref["loglik"]) < TOL and rel(got["fixed"][0, p], ref["fixed"]) < TOL
        assert rel(got["logdens"][0, p], ref["logdens"]) < TOL
    assert got["noos"][0, 3] == 2 * 1501 and np.isinf(got["fixed"][0, 2]) and np.isinf(got["fixed"][0, 4])
    eng.close()


@pytest.mark.parametrize("n_obs,p,ctas", [(1, 1, 1), (2, 2, 1), (33, 2, 2), (1500, 3, 0), (4099, 5, 4), (20001, 8, 8)])
def test_linreg_trajectory_parity(oracle, datasets, n_obs, p, ctas):
    spec = datasets.linear_regression(n_obs, p)
    rng = np.random.default_rng(n_obs)
    eng, ref = make_pair(oracle, spec, 6, seed=9 + n_obs, ctas=ctas, init="COPIES")
    start_from(eng, ref, [[v * (1.0 + 0.05 * rng.normal()) for v in spec["truth"]] for _ in range(6)])
    assert_trajectory_parity(eng, ref, 8)
    eng.close()
    if n_obs <= 1500:
        eng, ref = make_pair(oracle, spec, 6, seed=2, ctas=ctas)     # prior draws (FORWARD)
        assert_trajectory_parity(eng, ref, 6)
        eng.close()


def test_linreg_options_scm_rounds(oracle, datasets):
    from blangsdk_b200.engine import PT, Engine
    spec = datasets.linear_regression(400, 3)
    for opts in (dict(usePriorSamples=False), dict(reversible=True), dict(annealSupport=False, nPassesPerScan=1.5)):
        eng, ref = make_pair(oracle, spec, 5, seed=3, **opts)
        assert_trajectory_parity(eng, ref, 6)
        eng.close()
    eng = Engine(nChains=5, random=41)
    eng.set_model(spec)
    eng.initialize("SCM", nParticles=40, threshold=0.9)
    ref = oracle.PT(oracle.Model(spec), nChains=5, seed=41)
    rs, gs = ref.initialize_scm(nParticles=40, threshold=0.9), eng.scm_summary()
    assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
    assert_trajectory_parity(eng, ref, 5)
    eng.close()
    pt = PT(nChains=5, nScans=100, adaptFraction=0.5, random=13, initialization="FORWARD")
    pt.setSampledModel(spec)
    res = pt.performInference()
    ref = oracle.PT(oracle.Model(spec), nChains=5, seed=13)
    ref.initialize()
    rres = ref.perform_inference(100, 0.5)
    assert np.array_equal(res["swapIndicators"][:, 0, :], rres["indicators"])
    assert rel(res["energy"][:, 0, :], rres["energies"]) < TOL and rel(res["ladder"][0], ref.get_ladder()) < TOL
    pt.engine.close()


def test_linreg_posterior_against_least_squares(tmp_path, datasets):
    """the beta = 1 chain concentrates on the least-squares fit (flat-ish priors, 5 000 observations)"""
    import csv
    from blangsdk_b200.engine import PT
    spec = datasets.linear_regression(5000, 3)
    pt = PT(nChains=8, nScans=600, adaptFraction=0.5, random=4, initialization="SCM")
    pt.setSampledModel(spec)
    pt.performInference(results=str(tmp_path))
    rows = list(csv.DictReader(open(tmp_path / "samples" / "coefficients.csv")))
    w = np.array([[float(r["value"]) for r in rows if r["index_0"] == str(j)] for j in range(3)])
    sig = np.array([float(r["value"]) for r in csv.DictReader(open(tmp_path / "samples" / "sigma.csv"))])
    ols, resid, _, _ = np.linalg.lstsq(spec["x"], spec["y"], rcond=None)
    h = w.shape[1] // 2
    se = math.sqrt(resid[0] / 5000) / math.sqrt(5000)
    assert np.all(np.abs(w[:, h:].mean(axis=1) - ols) < 6 * se)
    assert abs(sig[h:].mean() - math.sqrt(resid[0] / 5000)) < 0.05
    pt.engine.close()
