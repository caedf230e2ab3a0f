"""-m gpu parity tests for family 7: F/HierarchicalModel.bl (the rocket-launch example of the Blang documentation),
not marginalised -- a, b ~ Exponential; p_g ~ Beta(a, b); failures_g ~ Binomial(launches_g, p_g)."""
import math

import numpy as np
import pytest

from gpu_util import assert_trajectory_parity, make_pair, rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("kind", ["regular", "oos"])
def test_hierarchical_log_density_parity(oracle, datasets, kind):
    from blangsdk_b200.engine import Engine
    spec = datasets.hierarchical_rockets(9)
    n = 6
    eng = Engine(nChains=n, random=1)
    eng.set_model(spec)
    m = oracle.Model(spec)
    assert (eng.n_reals, eng.n_samplers) == (m.n_reals, m.n_samplers) == (11, 11)
    rng = np.random.default_rng(3)
    states = []
    for p in range(n):
        st = [rng.gamma(2.0, 1.0), rng.gamma(2.0, 2.0)] + list(rng.uniform(0.02, 0.98, 9))
        if kind == "oos":
            j = int(rng.integers(0, 11))
            st[j] = -0.3 if j < 2 else 1.4      # a, b <= 0 or a probability outside (0, 1)
        states.append(st)
        eng.set_state(p, st, None)
    got = eng.log_density()
    ladder = eng.get_ladder()[0]
    for p in range(n):
        ref = m.log_density(states[p], None, beta=ladder[p])
        assert got["noos"][0, p] == ref["noos"]
        assert rel(got["loglik"][0, p], ref["loglik"]) < TOL and rel(got["fixed"][0, p], ref["fixed"]) < TOL
        assert rel(got["logdens"][0, p], ref["logdens"]) < TOL
    if kind == "oos":
        assert np.all(np.isinf(got["fixed"][0]))
    eng.close()


@pytest.mark.parametrize("groups", [1, 5, 31, 62])
def test_hierarchical_trajectory_parity(oracle, datasets, groups):
    spec = datasets.hierarchical_rockets(groups)
    for init in ("FORWARD", "COPIES"):
        eng, ref = make_pair(oracle, spec, 6, seed=17 + groups, init=init)
        assert_trajectory_parity(eng, ref, 8)
        eng.close()
    for opts in (dict(usePriorSamples=False), dict(reversible=True), dict(annealSupport=False, nPassesPerScan=0.7)):
        eng, ref = make_pair(oracle, spec, 5, seed=3, **opts)
        assert_trajectory_parity(eng, ref, 6)
        eng.close()


def test_hierarchical_scm_and_fused_replicas(oracle, datasets, monkeypatch):
    from blangsdk_b200.engine import Engine
    spec = datasets.hierarchical_rockets(7)
    N = 5
    eng = Engine(nChains=N, random=41)
    eng.set_model(spec)
    eng.initialize("SCM", nParticles=40, threshold=0.9)
    ref = oracle.PT(oracle.Model(spec), nChains=N, seed=41)
    rs, gs = ref.initialize_scm(nParticles=40, threshold=0.9), eng.scm_summary()
    assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
    assert_trajectory_parity(eng, ref, 6)
    eng.close()
    R, S = 9, 10
    monkeypatch.setenv("BLANGPT_FUSED", "1")
    eng = Engine(nChains=N, nReplicas=R, random=19)
    eng.set_model(spec)
    eng.initialize("FORWARD")
    l0 = eng.counters()["launches"]
    out = eng.run_scans(S, want=("energy", "swapIndicators"))
    assert eng.counters()["launches"] - l0 == 1
    for r in range(R):
        rr = oracle.PT(oracle.Model(spec), nChains=N, seed=19, replica=r)
        rr.initialize()
        for s in range(S):
            e, ind = rr.scan()
            assert np.array_equal(out["swapIndicators"][s, r], ind)
            assert rel(out["energy"][s, r], e) < TOL
    eng.close()


def test_hierarchical_agrees_with_the_marginalised_family(tmp_path, datasets):
    """integrating p_g out of family 7 gives family 5 (beta-binomial): the posterior of (a, b) and the evidence must
    agree between two independent runs of the two families on the same data, within Monte Carlo error"""
    import csv
    from blangsdk_b200.engine import PT
    spec = datasets.hierarchical_rockets(10)
    marg = dict(family="betabinomial", hyper=dict(rate=1.0), y=spec["y"], ntrials=spec["ntrials"])
    stats = {}
    for name, s in (("hier", spec), ("marg", marg)):
        pt = PT(nChains=10, nScans=4000, adaptFraction=0.5, random=7, initialization="SCM")
        pt.setSampledModel(s)
        pt.performInference(results=str(tmp_path / name))
        a_file, b_file = ("a.csv", "b.csv") if name == "hier" else ("alpha.csv", "beta.csv")
        a = np.array([float(r["value"]) for r in csv.DictReader(open(tmp_path / name / "samples" / a_file))])
        b = np.array([float(r["value"]) for r in csv.DictReader(open(tmp_path / name / "samples" / b_file))])
        z = float(list(csv.DictReader(open(tmp_path / name / "logNormalizationEstimate.csv")))[-1]["value"])
        h = len(a) // 2
        stats[name] = (a[h:].mean(), b[h:].mean(), np.log(a[h:]).std(), z)
        if name == "hier":
            fp = list(csv.DictReader(open(tmp_path / name / "samples" / "failureProbabilities.csv")))
            assert {r["rocketType"] for r in fp} == {str(g) for g in range(10)}
        pt.engine.close()
    (a1, b1, s1, z1), (a2, b2, s2, z2) = stats["hier"], stats["marg"]
    assert abs(math.log(a1 / a2)) < 0.25 and abs(math.log(b1 / b2)) < 0.25
    assert abs(z1 - z2) < 1.0
