"""-m gpu parity tests for family 9: F/SimpleHierarchicalModel.bl -- a, b ~ Exponential; p_g ~ Beta(a + 0.5, b + 0.5);
latent launch counts L_g ~ Poisson(2); failures_g ~ Binomial(1 + L_g, p_g).  The L_g are moved by IntSliceSampler
WITH stepping out (IntSliceSampler.java:53-132): the one sampler path no other family exercises on
the device.  The engine carries the integers as exact doubles after the reals; the oracle keeps them as ints."""
import math

import numpy as np
import pytest

from gpu_util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _split(state, G):
    s = np.asarray(state)
    return s[:G + 2], s[G + 2:].astype(np.int64)


def _pair(oracle, spec, n_chains, seed, init="FORWARD", **opts):
    from blangsdk_b200.engine import Engine
    eng = Engine(nChains=n_chains, random=seed, **opts)
    eng.set_model(spec)
    ref = oracle.PT(oracle.Model(spec), nChains=n_chains, seed=seed, nPassesPerScan=opts.get("nPassesPerScan", 3.0),
                    usePriorSamples=opts.get("usePriorSamples", True), reversible=opts.get("reversible", False),
                    annealSupport=opts.get("annealSupport", True))
    if init == "FORWARD":
        eng.initialize("FORWARD")
        ref.initialize(oracle.INIT_FORWARD)
    return eng, ref


def _feasible_start(eng, ref, spec, seed, replicas=1):
    """every chain from a state with 1 + L_g >= failures_g: at beta = 1 an infeasible state has density zero, the
    integer slice sampler then finds no acceptable point and the reference throws (see the test below)"""
    rng = np.random.default_rng(seed)
    G = len(spec["y"])
    for p in range(eng.N):
        re = [float(rng.gamma(2.0, 0.5)) + 0.2, float(rng.gamma(2.0, 0.5)) + 0.2] + list(rng.uniform(0.1, 0.9, G))
        it = [int(v) for v in np.maximum(spec["y"] - 1, 0) + rng.poisson(1.0, G)]
        for r in range(replicas):
            eng.set_state(p, re + [float(v) for v in it], None, replica=r)
        if ref is not None:
            ref.set_state(p, re, it)
    eng.initialize("COPIES")


def _assert_parity(eng, ref, scans, G):
    out = eng.run_scans(scans, want=("energy", "swapIndicators", "samples"))
    for s in range(scans):
        e, ind = ref.scan()
        assert np.array_equal(out["swapIndicators"][s, 0], ind), "swap decisions differ at scan %d" % s
        assert rel(out["energy"][s, 0], e) < TOL
    assert eng.counters()["evals"] == ref.n_evals(), "number of log-density evaluations differs"
    for p in range(eng.N):
        st, _ = eng.get_state(p)
        re, it = _split(st, G)
        rr, ri = ref.get_state(p)
        assert rel(re, rr) < TOL and np.array_equal(it, ri)


@pytest.mark.parametrize("kind", ["regular", "oos"])
def test_rockets_latent_log_density_parity(oracle, datasets, kind):
    from blangsdk_b200.engine import Engine
    G = 7
    spec = datasets.rockets_latent(G)
    n = 6
    eng = Engine(nChains=n, random=1)
    eng.set_model(spec)
    m = oracle.Model(spec)
    assert eng.n_reals == 2 * G + 2 and (m.n_reals, m.n_ints, m.n_samplers) == (G + 2, G, 2 * G + 2) and eng.n_samplers == 2 * G + 2
    rng = np.random.default_rng(3)
    states = []
    for p in range(n):
        re = [rng.gamma(2.0, 1.0), rng.gamma(2.0, 2.0)] + list(rng.uniform(0.02, 0.98, G))
        it = list(np.maximum(spec["y"] - 1, 0) + rng.poisson(1.5, G))       # 1 + L >= failures
        if kind == "oos":
            j = int(rng.integers(0, 3))
            if j == 0:
                it[int(rng.integers(0, G))] = -2                             # negative count: Poisson and Binomial blocks
            elif j == 1:
                g = int(np.argmax(spec["y"])); it[g] = max(0, int(spec["y"][g]) - 2) if spec["y"][g] >= 2 else -1   # failures > trials
            else:
                re[0] = -0.7                                                 # a + 0.5 <= 0
        states.append((re, it))
        eng.set_state(p, re + [float(v) for v in it], None)
    got = eng.log_density()
    ladder = eng.get_ladder()[0]
    for p in range(n):
        ref = m.log_density(states[p][0], states[p][1], beta=ladder[p])
        assert got["noos"][0, p] == ref["noos"]
        assert rel(got["loglik"][0, p], ref["loglik"]) < TOL and rel(got["fixed"][0, p], ref["fixed"]) < TOL
        assert rel(got["logdens"][0, p], ref["logdens"]) < TOL
    if kind == "oos":
        assert got["noos"].sum() > 0 or np.any(np.isinf(got["fixed"]))
    eng.close()


@pytest.mark.parametrize("groups", [1, 4, 13, 31])
def test_rockets_latent_trajectory_parity(oracle, datasets, groups):
    """real slice on a, b, p_g and the integer slice sampler with stepping out on every L_g, draw for draw"""
    spec = datasets.rockets_latent(groups)
    eng, ref = _pair(oracle, spec, 6, seed=23 + groups, init="COPIES")
    _feasible_start(eng, ref, spec, groups)
    _assert_parity(eng, ref, 10, groups)
    assert ref.error() == 0
    eng.close()
    for opts in (dict(usePriorSamples=False), dict(reversible=True), dict(annealSupport=False, nPassesPerScan=0.7)):
        eng, ref = _pair(oracle, spec, 5, seed=3, init="COPIES", **opts)
        _feasible_start(eng, ref, spec, 100 + groups)
        _assert_parity(eng, ref, 8, groups)
        assert ref.error() == 0
        eng.close()


def test_rockets_latent_infeasible_start_is_an_error(oracle, datasets):
    """L = 0 everywhere (COPIES) with groups of >= 2 failures: at beta = 1 the state has density zero, the integer
    sampler's window [x0 - U, x0 - U + 10) may hold no feasible count, stepping out does not start (both ends are
    -inf < the -1e100 slice height) and the reference's shrinkage ends in discreteUniform throwing
    "Invalid arguments".  The oracle records it; the engine reports it instead of hanging."""
    from blangsdk_b200.engine import BlangPTError
    spec = datasets.rockets_latent(13)
    assert spec["y"].max() >= 3
    eng, ref = _pair(oracle, spec, 6, seed=36, init="COPIES")
    for _ in range(10):
        ref.scan()
    assert ref.error() & 4
    with pytest.raises(BlangPTError, match="Invalid arguments"):
        eng.run_scans(10, want=("energy",))
        eng.counters()
    eng.close()


def test_rockets_latent_scm_and_fused_replicas(oracle, datasets, monkeypatch):
    from blangsdk_b200.engine import Engine
    G, N = 5, 5
    spec = datasets.rockets_latent(G)
    eng = Engine(nChains=N, random=41)
    eng.set_model(spec)
    eng.initialize("SCM", nParticles=40, threshold=0.9)
    ref = oracle.PT(oracle.Model(spec), nChains=N, seed=41)
    rs, gs = ref.initialize_scm(nParticles=40, threshold=0.9), eng.scm_summary()
    assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
    if ref.error() == 0:
        _assert_parity(eng, ref, 6, G)
    eng.close()
    R, S = 9, 10
    monkeypatch.setenv("BLANGPT_FUSED", "1")
    eng = Engine(nChains=N, nReplicas=R, random=19)
    eng.set_model(spec)
    _feasible_start(eng, None, spec, 7, replicas=R)
    l0 = eng.counters()["launches"]
    out = eng.run_scans(S, want=("energy", "swapIndicators"))
    assert eng.counters()["launches"] - l0 == 1
    for r in range(R):
        rr = oracle.PT(oracle.Model(spec), nChains=N, seed=19, replica=r)
        rng = np.random.default_rng(7)
        for p in range(N):
            re = [float(rng.gamma(2.0, 0.5)) + 0.2, float(rng.gamma(2.0, 0.5)) + 0.2] + list(rng.uniform(0.1, 0.9, G))
            it = [int(v) for v in np.maximum(spec["y"] - 1, 0) + rng.poisson(1.0, G)]
            rr.set_state(p, re, it)
        for s in range(S):
            e, ind = rr.scan()
            assert np.array_equal(out["swapIndicators"][s, r], ind)
            assert rel(out["energy"][s, r], e) < TOL
    eng.close()


def test_rockets_latent_counts_posterior(tmp_path, datasets):
    """end to end, tables on disk: the conditional law of a latent count given its group's failures k and probability
    p is known by enumeration, P(L | k, p) ~ Poisson(L; mean) C(1 + L, k) p^k (1 - p)^(1 + L - k); averaged over the
    run's own samples of p it must reproduce the run's mean of L"""
    import csv
    from blangsdk_b200.engine import PT
    spec = datasets.rockets_latent(6)
    pt = PT(nChains=8, nScans=3000, adaptFraction=0.5, random=5, initialization="SCM")
    pt.setSampledModel(spec)
    pt.performInference(results=str(tmp_path))
    rows = list(csv.DictReader(open(tmp_path / "samples" / "numberOfLaunches.csv")))
    probs = list(csv.DictReader(open(tmp_path / "samples" / "failureProbabilities.csv")))
    G = 6
    L = np.array([[float(r["value"]) for r in rows if r["rocketType"] == str(g)] for g in range(G)])
    P = np.array([[float(r["value"]) for r in probs if r["rocketType"] == str(g)] for g in range(G)])
    assert np.all(L == np.round(L)) and np.all(1 + L >= spec["y"][:, None])       # integers, always >= failures - 1
    h = L.shape[1] // 2
    from scipy.stats import binom, poisson
    Ls = np.arange(0, 120)
    for g in range(G):
        k = int(spec["y"][g])
        cond = []
        for p in P[g, h::15]:
            w = poisson.pmf(Ls, 2.0) * binom.pmf(k, 1 + Ls, p)
            cond.append((w * Ls).sum() / w.sum())
        assert abs(L[g, h:].mean() - np.mean(cond)) < 0.2, (g, k, L[g, h:].mean(), np.mean(cond))
    pt.engine.close()


def test_scm_zero_over_zero_ratio_is_the_references_error(oracle, datasets):
    """After T/runtime/TestSampledModel.java: with the support not annealed, prior draws whose launch counts cannot
    explain the failures have density zero at every beta > 0, the SCM weight update meets 0/0 and the reference throws
    SampledModel.INVALID_LOG_RATIO; so do the oracle and the engine."""
    from blangsdk_b200.engine import BlangPTError, Engine
    spec = datasets.rockets_latent(13)
    ref = oracle.PT(oracle.Model(spec), nChains=4, seed=1, annealSupport=False)
    with pytest.raises(RuntimeError, match="Invalid logDensity ratio"):
        ref.initialize_scm(nParticles=8, threshold=0.9)
    eng = Engine(nChains=4, random=1, annealSupport=False)
    eng.set_model(spec)
    with pytest.raises(BlangPTError, match="Invalid logDensity ratio"):
        eng.initialize("SCM", nParticles=8, threshold=0.9)
    eng.close()
