"""-m gpu parity tests for family 6 (i.i.d. observations from Poisson / Binomial / NegativeBinomial / StudentT /
Geometric / Laplace with one prior per parameter).  Same bar as test_gpu_parity.py: densities within
1e-9 relative, out-of-support counts, swap decisions and evaluation counts identical."""
import math

import numpy as np
import pytest

from gpu_util import assert_trajectory_parity, make_pair, rel, start_from

pytestmark = pytest.mark.gpu
TOL = 1e-9
DISTS = ["poisson", "binomial", "negbinomial", "studentt", "geometric", "laplace", "exponential", "gamma", "weibull",
         "gumbel", "logisticdist", "halfstudentt", "loglogistic", "gompertz", "yulesimon", "uniformmax"]


def _state(rng, dist, kind):
    """a parameter vector inside the support ("regular") or with one parameter outside ("oos")"""
    th = {"poisson": [rng.gamma(2.0, 2.0)], "binomial": [rng.uniform(0.05, 0.95)],
          "negbinomial": [rng.gamma(2.0, 1.5), rng.uniform(0.05, 0.95)],
          "studentt": [rng.gamma(2.0, 2.0), rng.normal(1.0, 2.0), rng.gamma(2.0, 1.0)],
          "geometric": [rng.uniform(0.05, 0.95)], "laplace": [rng.normal(0.0, 2.0), rng.gamma(2.0, 1.0)],
          "exponential": [rng.gamma(2.0, 1.0)], "gamma": [rng.gamma(2.0, 1.0), rng.gamma(2.0, 1.0)],
          "weibull": [rng.gamma(2.0, 1.0), rng.gamma(2.0, 1.0)], "gumbel": [rng.normal(0.0, 2.0), rng.gamma(2.0, 1.0)],
          "logisticdist": [rng.normal(0.0, 2.0), rng.gamma(2.0, 1.0)], "halfstudentt": [rng.gamma(2.0, 2.0), rng.gamma(2.0, 1.0)],
          "loglogistic": [rng.gamma(2.0, 1.0), rng.gamma(2.0, 1.0)], "gompertz": [rng.gamma(2.0, 0.5), rng.gamma(2.0, 1.0)],
          "yulesimon": [rng.gamma(2.0, 1.0)], "uniformmax": [rng.uniform(1.0, 6.0)]}[dist]
    if kind == "oos":
        j = {"poisson": 0, "binomial": 0, "negbinomial": int(rng.integers(0, 2)), "studentt": int(rng.choice([0, 2])),
             "geometric": 0, "laplace": 1, "exponential": 0, "gamma": int(rng.integers(0, 2)),
             "weibull": int(rng.integers(0, 2)), "gumbel": 1, "logisticdist": 1, "halfstudentt": int(rng.integers(0, 2)),
             "loglogistic": int(rng.integers(0, 2)), "gompertz": int(rng.integers(0, 2)), "yulesimon": 0, "uniformmax": 0}[dist]
        th[j] = -abs(th[j]) - 0.1
    return th


@pytest.mark.parametrize("dist", DISTS)
@pytest.mark.parametrize("kind", ["regular", "oos"])
def test_iid_log_density_parity(oracle, datasets, dist, kind):
    from blangsdk_b200.engine import Engine
    spec = datasets.iid_observations(dist, 777)
    n = 6
    eng = Engine(nChains=n, random=1)
    eng.set_model(spec)
    m = oracle.Model(spec)
    assert (eng.n_reals, eng.n_samplers) == (m.n_reals, m.n_samplers)
    rng = np.random.default_rng(7)
    states = [_state(rng, dist, kind) for _ in range(n)]
    for p in range(n):
        eng.set_state(p, states[p], None)
    got = eng.log_density()
    ladder = eng.get_ladder()[0]
    for p in range(n):
        ref = m.log_density(states[p], None, beta=ladder[p])
        assert got["noos"][0, p] == ref["noos"]
        assert rel(got["loglik"][0, p], ref["loglik"]) < TOL
        assert rel(got["fixed"][0, p], ref["fixed"]) < TOL
        if math.isinf(ref["logdens"]) or math.isnan(ref["logdens"]):
            assert rel(got["logdens"][0, p], ref["logdens"]) == 0.0
        else:
            assert abs(got["logdens"][0, p] - ref["logdens"]) <= TOL * max(1.0, abs(ref["logdens"]))
    if kind == "oos":
        assert got["noos"].sum() >= 777 * n
    eng.close()


@pytest.mark.parametrize("dist", DISTS)
def test_iid_trajectory_parity(oracle, datasets, dist):
    """draw-for-draw: swap decisions, evaluation counts and states agree with the oracle, for several cluster sizes
    and observation counts (ragged tails included)"""
    rng = np.random.default_rng(11)
    for n_obs, ctas in ((1, 1), (33, 1), (1501, 2), (4097, 4)):
        spec = datasets.iid_observations(dist, n_obs)
        # chains start near the generating parameters: from a wide prior draw the log-likelihood of thousands of
        # observations under exp/pow laws is ~1e12 (or overflows), where 1e-16 relative rounding differences between
        # libm and CUDA exceed the slice sampler's O(1) comparisons and parity stops being a meaningful test
        eng, ref = make_pair(oracle, spec, 7, seed=100 + n_obs, ctas=ctas, init="COPIES")
        start_from(eng, ref, [[v * (1.0 + 0.1 * rng.normal()) for v in spec["truth"]] for _ in range(7)])
        assert_trajectory_parity(eng, ref, 8)
        eng.close()
    if dist in ("poisson", "binomial", "negbinomial", "studentt", "geometric", "laplace", "exponential", "gamma"):
        eng, ref = make_pair(oracle, datasets.iid_observations(dist, 500), 7, seed=5)      # prior draws (FORWARD)
        assert_trajectory_parity(eng, ref, 8)
        eng.close()


@pytest.mark.parametrize("opts", [dict(usePriorSamples=False), dict(reversible=True), dict(annealSupport=False),
                                  dict(nPassesPerScan=1.5)])
def test_iid_engine_options(oracle, datasets, opts):
    for dist in ("negbinomial", "studentt"):
        eng, ref = make_pair(oracle, datasets.iid_observations(dist, 400), 6, seed=77, **opts)
        assert_trajectory_parity(eng, ref, 8)
        eng.close()


def test_iid_start_outside_the_support(oracle, datasets):
    """latent parameters start at 0.0 (COPIES): p = 0 is inside the Uniform(0,1) prior but outside the law's support
    (every observation's factor is -inf); the softened support brings the chains back, as in the oracle"""
    spec = datasets.iid_observations("geometric", 300)
    eng, ref = make_pair(oracle, spec, 5, seed=3, init="COPIES")
    d = eng.log_density()
    assert np.all(d["noos"][0] == 300) and np.all(d["fixed"][0] == 0.0)
    assert_trajectory_parity(eng, ref, 10)
    assert np.all(eng.log_density()["noos"][0, :-1] == 0)
    eng.close()


def test_iid_scm_rounds_and_replicas(oracle, datasets, monkeypatch):
    from blangsdk_b200.engine import Engine
    spec = datasets.iid_observations("studentt", 250)
    N = 5
    # SCM initialisation (PT's default)
    eng = Engine(nChains=N, random=41)
    eng.set_model(spec)
    eng.initialize("SCM", nParticles=40, threshold=0.9)
    ref = oracle.PT(oracle.Model(spec), nChains=N, seed=41)
    rs, gs = ref.initialize_scm(nParticles=40, threshold=0.9), eng.scm_summary()
    assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
    assert_trajectory_parity(eng, ref, 6)
    eng.close()
    # rounds with schedule adaptation (PT.performInference)
    from blangsdk_b200.engine import PT
    pt = PT(nChains=N, nScans=120, adaptFraction=0.5, random=13, initialization="FORWARD")
    pt.setSampledModel(spec)
    res = pt.performInference()
    ref = oracle.PT(oracle.Model(spec), nChains=N, seed=13)
    ref.initialize()
    rres = ref.perform_inference(120, 0.5)
    assert [r["nScans"] for r in res["rounds"]] == [r["nScans"] for r in rres["rounds"]]
    assert np.array_equal(res["swapIndicators"][:, 0, :], rres["indicators"])
    assert rel(res["energy"][:, 0, :], rres["energies"]) < TOL
    assert rel(res["samples"][:, 0, :], rres["samples"]) < TOL
    assert rel(res["ladder"][0], ref.get_ladder()) < TOL
    for a, b in zip(res["rounds"], rres["rounds"]):
        assert int(a["roundTrips"][0]) == b["roundTrips"]
        assert rel(a["Lambda"][0], b["Lambda"]) < TOL and rel(a["logZ"][0], b["logZ"]) < TOL
    pt.engine.close()
    # many replicas in the fused kernel, each == its own oracle run
    R, S = 9, 12
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


def test_iid_error_paths(datasets):
    from blangsdk_b200.engine import BlangPTError, Engine
    spec = datasets.iid_observations("binomial", 50)
    bad = dict(spec, hyper=dict(dist="binomial", priors=[("beta", 2.0, 3.0), ("beta", 1.0, 1.0)]))
    with pytest.raises(BlangPTError):
        Engine(nChains=2).set_model(bad)                       # wrong number of priors
    with pytest.raises(BlangPTError):
        Engine(nChains=2).set_model(dict(spec, hyper=dict(dist="cauchy", priors=[])))
    with pytest.raises(BlangPTError):
        Engine(nChains=2).set_model(dict(spec, hyper=dict(dist="binomial", priors=[("pareto", 1.0)])))


def test_iid_posterior_against_conjugate_answers(tmp_path, datasets):
    """a full run through the PT mirror, tables on disk; the beta = 1 chain's posterior means match the conjugate
    closed forms (Gamma-Poisson, Beta-Binomial) within Monte Carlo error, and so does log Z for the Poisson model"""
    import csv
    from blangsdk_b200.engine import PT
    # Poisson with Exponential(0.1) = Gamma(1, 0.1) prior: posterior Gamma(1 + sum y, 0.1 + n)
    spec = datasets.iid_observations("poisson", 400)
    y = spec["y"]
    pt = PT(nChains=8, nScans=1500, adaptFraction=0.5, random=5, initialization="SCM")
    pt.setSampledModel(spec)
    res = pt.performInference(results=str(tmp_path / "poisson"))
    mean = np.array([float(r["value"]) for r in csv.DictReader(open(tmp_path / "poisson" / "samples" / "mean.csv"))])
    post_mean = (1 + y.sum()) / (0.1 + len(y))
    post_sd = math.sqrt(1 + y.sum()) / (0.1 + len(y))
    tail = mean[len(mean) // 2:]
    assert abs(tail.mean() - post_mean) < 5 * post_sd / math.sqrt(len(tail) / 4.0)
    assert 0.7 * post_sd < tail.std() < 1.4 * post_sd
    # marginal likelihood: log Z = log Gamma-Poisson evidence
    a0, b0, S, n = 1.0, 0.1, y.sum(), len(y)
    log_z = a0 * math.log(b0) - math.lgamma(a0) + math.lgamma(a0 + S) - (a0 + S) * math.log(b0 + n) - sum(math.lgamma(v + 1) for v in y)
    est = float(list(csv.DictReader(open(tmp_path / "poisson" / "logNormalizationEstimate.csv")))[-1]["value"])
    assert abs(est - log_z) < 1.0
    pt.engine.close()
    # Binomial with Beta(2, 3) prior: posterior Beta(2 + sum k, 3 + sum (n - k))
    spec = datasets.iid_observations("binomial", 200)
    k, tr = spec["y"], spec["trials"]
    pt = PT(nChains=8, nScans=1500, adaptFraction=0.5, random=6, initialization="SCM")
    pt.setSampledModel(spec)
    pt.performInference(results=str(tmp_path / "binomial"))
    p = np.array([float(r["value"]) for r in csv.DictReader(open(tmp_path / "binomial" / "samples" / "probabilityOfSuccess.csv"))])
    a, b = 2 + k.sum(), 3 + (tr - k).sum()
    sd = math.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)))
    tail = p[len(p) // 2:]
    assert abs(tail.mean() - a / (a + b)) < 5 * sd / math.sqrt(len(tail) / 4.0)
    pt.engine.close()


def test_doomsday_on_the_device(oracle, datasets):
    """F/Doomsday.bl with the arguments of the reference's Getting Started transcript and of T/TestEndToEnd.xtend:38-47
    (rate 1.0, y 1.2, z latent): draw-for-draw parity with the oracle from PT's default SCM initialisation, the same SCM
    evidence estimate, and the device's stepping-stone log Z against the closed form log E1(1.2) and the number the JVM
    printed (-1.8657991502743467, GettingStarted.xtend:255)."""
    import math
    from scipy.special import exp1
    from blangsdk_b200.engine import Engine, PT
    spec = datasets.doomsday(1.2, 1.0)
    exact, jvm = math.log(exp1(1.2)), -1.8657991502743467
    eng = Engine(nChains=8, random=77)
    eng.set_model(spec)
    eng.initialize("SCM", nParticles=200)
    ref = oracle.PT(oracle.Model(spec), nChains=8, seed=77)
    rs, gs = ref.initialize_scm(nParticles=200), eng.scm_summary()
    assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
    assert abs(gs["logNormEstimate"] - rs["logNormEstimate"]) < 1e-9
    assert_trajectory_parity(eng, ref, 25)
    eng.close()
    est = []
    for seed in range(10):
        pt = PT(nChains=8, nScans=1000, adaptFraction=0.5, random=900 + seed, initialization="FORWARD")
        pt.setSampledModel(spec)
        res = pt.performInference(want=())
        est.append(float(res["rounds"][-1]["logZ"][0]))
        pt.engine.close()
    est = np.array(est)
    se = est.std(ddof=1) / math.sqrt(len(est))
    assert abs(est.mean() - exact) < 4 * se + 1e-3, (est.mean(), exact, se)
    assert abs(est.mean() - jvm) < abs(jvm - exact) + 4 * se + 1e-3
