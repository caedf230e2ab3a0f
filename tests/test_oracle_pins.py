"""Pins the CPU oracle on the known answers the reference's own tests hold for the NRPT path
(SURVEY.md 8(c)).  No JVM exists here, so these are the anchors -- the oracle header says
"parity unpinned" with respect to JVM outputs."""
import math

import numpy as np
import pytest


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_ladders(oracle):
    # T/TestLadders.java:44-81
    assert list(oracle.ladder_equally_spaced(1)) == [1.0]
    assert list(oracle.ladder_equally_spaced(2)) == [1.0, 0.0]
    lad = oracle.ladder_equally_spaced(5)
    assert len(lad) == 5 and lad[0] == 1.0 and lad[-1] == 0.0
    assert np.allclose(lad, [1.0, 0.75, 0.5, 0.25, 0.0])


def test_rounds_worked_example(oracle):
    # PT.rounds, PT.java:239-265: n=1000, f=0.5 -> 1001 scans
    r = oracle.rounds(1000, 0.5)
    assert [x[0] for x in r] == [2, 4, 8, 16, 31, 63, 126, 251, 500]
    assert [x[1] for x in r] == [True] * 8 + [False]
    assert sum(x[0] for x in r) == 1001
    assert oracle.rounds(10, 0.0) == [(10, False)]


def _simpson(f, a, b, n=20000):
    x = np.linspace(a, b, n + 1)
    y = f(x)
    h = (b - a) / n
    return h / 3 * (y[0] + y[-1] + 4 * y[1:-1:2].sum() + 2 * y[2:-1:2].sum())


def test_density_normalisation(oracle, datasets):
    """T/TestSDKNormalizations.java:20-45: Normal / Exponential(Gamma) integrate to 1 +- 1e-6; the
    beta-binomial and Bernoulli(logistic) pmfs sum to 1."""
    # Normal likelihood: one observation y, integrate exp(loglik) over y
    def norm_lik(ys, mu=0.7, s2=1.9):
        out = []
        for y in ys:
            m = oracle.Model(dict(family="normal", hyper=dict(m0=0.0, v0=100.0, rate=0.1), y=np.array([y])))
            out.append(math.exp(m.log_density([mu, s2])["loglik"]))
        return np.array(out)
    assert abs(_simpson(norm_lik, 0.7 - 14, 0.7 + 14, 400) - 1.0) < 1e-6
    # priors of the Normal family: Normal(0,100) on mu  x  Exponential(0.1) on s2
    m = oracle.Model(datasets.normal_meanvar(n=1))
    base = m.log_density([0.3, 2.0])
    def prior_mu(xs):
        return np.array([math.exp(m.log_density([x, 2.0])["fixed"] - base["fixed"]) for x in xs])
    def prior_s2(xs):
        return np.array([math.exp(m.log_density([0.3, x])["fixed"] - base["fixed"]) if x > 0 else 0.0 for x in xs])
    # fixed(mu, s2) = logN(mu) + logExp(s2): ratios integrate to 1/pdf at the base point
    pdf_mu = math.exp(-0.5 * math.log(2 * math.pi * 100.0) - 0.5 * 0.3 ** 2 / 100.0)
    pdf_s2 = 0.1 * math.exp(-0.1 * 2.0)
    assert abs(_simpson(prior_mu, -150, 150, 2000) * pdf_mu - 1.0) < 1e-6
    assert abs(_simpson(prior_s2, 1e-12, 400.0, 8000) * pdf_s2 - 1.0) < 1e-6
    assert abs(math.exp(base["fixed"]) - pdf_mu * pdf_s2) < 1e-12
    # beta-binomial pmf sums to one
    tot = 0.0
    for y in range(0, 13):
        mb = oracle.Model(dict(family="betabinomial", hyper=dict(rate=0.1), y=np.array([y]), ntrials=np.array([12])))
        tot += math.exp(mb.log_density([1.7, 3.2])["loglik"])
    assert abs(tot - 1.0) < 1e-9
    # Bernoulli(logistic(eta)) sums to one
    x = np.array([[1.0, -0.3]])
    p = [math.exp(oracle.Model(dict(family="logistic", hyper=dict(v0=10.0), x=x, y=np.array([yy]))).log_density([0.4, 1.1])["loglik"])
         for yy in (0, 1)]
    assert abs(sum(p) - 1.0) < 1e-12


def _iid_loglik(oracle, dist, y, theta, trials=None):
    priors = {"poisson": [("exponential", 0.1)], "binomial": [("beta", 2.0, 3.0)],
              "negbinomial": [("gamma", 2.0, 0.5), ("uniform", 0.0, 1.0)],
              "studentt": [("exponential", 0.1), ("normal", 0.0, 100.0), ("exponential", 0.5)],
              "geometric": [("uniform", 0.0, 1.0)], "laplace": [("normal", 0.0, 25.0), ("gamma", 2.0, 1.0)]}.get(dist)
    if priors is None:
        n_par = dict(exponential=1, gamma=2, weibull=2, gumbel=2, logisticdist=2, halfstudentt=2, loglogistic=2,
                     gompertz=2, yulesimon=1, uniformmax=1)[dist]
        priors = [("exponential", 0.5)] * n_par
    spec = dict(family="iid", hyper=dict(dist=dist, priors=priors), y=np.array([float(y)]),
                trials=None if trials is None else np.array([float(trials)]))
    return oracle.Model(spec).log_density(theta)


def test_iid_laws_normalise(oracle):
    """T/TestSDKNormalizations.java:20-45 for the laws of family 6: every pmf sums, every pdf integrates, to
    1 +- 1e-6 (the reference's own tolerance), priors included."""
    tot = sum(math.exp(_iid_loglik(oracle, "poisson", y, [3.7])["loglik"]) for y in range(0, 80))
    assert abs(tot - 1.0) < 1e-9
    tot = sum(math.exp(_iid_loglik(oracle, "binomial", y, [0.37], trials=17)["loglik"]) for y in range(0, 18))
    assert abs(tot - 1.0) < 1e-9
    tot = sum(math.exp(_iid_loglik(oracle, "negbinomial", y, [2.6, 0.45])["loglik"]) for y in range(0, 400))
    assert abs(tot - 1.0) < 1e-9
    tot = sum(math.exp(_iid_loglik(oracle, "geometric", y, [0.2])["loglik"]) for y in range(0, 400))
    assert abs(tot - 1.0) < 1e-9
    lap = lambda ys: np.array([math.exp(_iid_loglik(oracle, "laplace", y, [-0.5, 1.3])["loglik"]) for y in ys])
    assert abs(_simpson(lap, -0.5 - 60, -0.5, 3000) + _simpson(lap, -0.5, -0.5 + 60, 3000) - 1.0) < 1e-6
    # Student t: heavy tails, integrate on a finite window and add the exact tail mass of t_5 beyond 400 scale units
    stu = lambda ys: np.array([math.exp(_iid_loglik(oracle, "studentt", y, [5.0, 1.0, 2.0])["loglik"]) for y in ys])
    assert abs(_simpson(stu, 1.0 - 800.0, 1.0 + 800.0, 40000) - 1.0) < 1e-6
    # the other continuous laws of family 6: integrate exp(loglik) over the realization
    def pdf(dist, theta):
        return lambda ys: np.array([math.exp(_iid_loglik(oracle, dist, y, theta)["loglik"]) for y in ys])
    def on_log_scale(f):      # int f(y) dy = int f(e^u) e^u du: smooth even where f ~ y^(shape-1) at the origin
        return lambda us: f(np.exp(us)) * np.exp(us)
    assert abs(_simpson(pdf("exponential", [1.7]), 1e-12, 40.0, 4000) - 1.0) < 1e-6
    assert abs(_simpson(pdf("gamma", [2.5, 1.2]), 1e-12, 60.0, 6000) - 1.0) < 1e-6
    assert abs(_simpson(on_log_scale(pdf("weibull", [2.0, 1.5])), -40.0, math.log(40.0), 4000) - 1.0) < 1e-6
    assert abs(_simpson(pdf("gumbel", [0.5, 1.2]), -12.0, 45.0, 6000) - 1.0) < 1e-6
    assert abs(_simpson(pdf("logisticdist", [1.0, 0.8]), -30.0, 32.0, 6000) - 1.0) < 1e-6
    assert abs(_simpson(pdf("halfstudentt", [5.0, 1.5]), 0.0, 900.0, 40000) - 1.0) < 1e-6
    assert abs(_simpson(on_log_scale(pdf("loglogistic", [2.0, 3.0])), -30.0, 30.0, 4000) - 1.0) < 1e-6
    assert abs(_simpson(pdf("gompertz", [0.8, 1.5]), 0.0, 12.0, 4000) - 1.0) < 1e-6
    tot = sum(math.exp(_iid_loglik(oracle, "yulesimon", y, [2.5])["loglik"]) for y in range(0, 3000))
    assert abs(tot - 1.0) < 1e-6      # tail ~ Gamma(3.5) k^-2.5: 3000^-2.5 * 3.3 = 7e-9
    assert abs(_simpson(pdf("uniformmax", [2.3]), 0.0, 2.3, 200) - 1.0) < 1e-9      # ContinuousUniform(0, 2.3)
    assert _iid_loglik(oracle, "uniformmax", 2.4, [2.3])["noos"] == 1 and _iid_loglik(oracle, "uniformmax", 0.5, [-1.0])["noos"] == 2
    assert _iid_loglik(oracle, "gamma", 1.3, [-1.0, 2.0])["noos"] == 2 and _iid_loglik(oracle, "gamma", 1.3, [1.0, -2.0])["noos"] == 3
    assert _iid_loglik(oracle, "weibull", -1.0, [1.0, 2.0])["noos"] == 2
    # out-of-support parameters: every law that reads them is -inf, one count per observation and law
    assert _iid_loglik(oracle, "poisson", 3, [-1.0])["noos"] == 2
    assert _iid_loglik(oracle, "studentt", 0.3, [-1.0, 0.0, 1.0])["noos"] == 2      # nu: its own law and the joint one
    assert _iid_loglik(oracle, "studentt", 0.3, [3.0, 0.0, -1.0])["noos"] == 2      # sigma
    assert _iid_loglik(oracle, "negbinomial", 2, [1.5, 1.5])["noos"] == 1
    assert _iid_loglik(oracle, "binomial", 9, [0.5], trials=4)["noos"] == 2         # k > n: both laws
    # priors: Beta(2,3) and Uniform(0,1) integrate to one, Gamma(2, 0.5) too
    base = _iid_loglik(oracle, "binomial", 3, [0.4], trials=10)
    beta_pdf = lambda xs: np.array([math.exp(_iid_loglik(oracle, "binomial", 3, [x], trials=10)["fixed"]) for x in xs])
    assert abs(_simpson(beta_pdf, 1e-9, 1 - 1e-9, 2000) - 1.0) < 1e-6
    assert abs(math.exp(base["fixed"]) - 12.0 * 0.4 * 0.6 ** 2) < 1e-12              # Beta(2,3) pdf = 12 x (1-x)^2
    joint = lambda xs: np.array([math.exp(_iid_loglik(oracle, "negbinomial", 1, [x, 0.3])["fixed"]) for x in xs])
    assert abs(_simpson(joint, 1e-12, 120.0, 12000) - 1.0) < 1e-6                    # Gamma(2,0.5) x Uniform(0,1) at p = 0.3
    assert _iid_loglik(oracle, "negbinomial", 1, [1.0, 1.2])["fixed"] == -math.inf


def test_lngamma_choice(oracle):
    """libm lgamma (the documented choice) vs the Pike & Hill series bayonet is believed to use:
    the discrepancy class is < 3e-10 absolute, inside the 1e-9 relative bar for C5."""
    L = oracle.lib()
    xs = np.concatenate([np.linspace(0.01, 7, 400), np.linspace(7, 200, 400)])
    d = max(abs(L.orc_lngamma(float(x)) - L.orc_lngamma_pike_hill(float(x))) for x in xs)
    assert d < 3e-10, d
    assert L.orc_lngamma(1.0) == 0.0 and abs(L.orc_lngamma(5.0) - math.log(24.0)) < 1e-14
    assert abs(L.orc_logistic(0.0) - 0.5) == 0 and abs(L.orc_log_add(math.log(2), math.log(3)) - math.log(5)) < 1e-15


def test_annealing_semantics(oracle, datasets):
    """T/TestEndToEnd.xtend:169-186 pins beta*loglik annealing to 1e-10; SampledModel.java:161-176."""
    m = oracle.Model(datasets.normal_meanvar(n=50))
    rng = np.random.default_rng(0)
    for _ in range(20):
        st = [rng.normal(), rng.gamma(2.0)]
        beta = rng.random()
        d = m.log_density(st, beta=beta)
        assert abs(d["logdens"] - (d["fixed"] + beta * d["loglik"])) < 1e-10 * max(1, abs(d["logdens"]))
        assert d["noos"] == 0
    # out of support: variance <= 0 makes 2n annealed factors -inf, and the fixed Gamma factors too
    d = m.log_density([0.0, -1.0], beta=0.5)
    assert d["noos"] == 100 and d["logdens"] == -math.inf and d["fixed"] == -math.inf
    # softened -inf: logistic row with p == 1.0 and y == 0 (prior support is wider than the likelihood's)
    x = np.array([[1.0]]); y = np.array([0])
    ml = oracle.Model(dict(family="logistic", hyper=dict(v0=10.0), x=x, y=y))
    d = ml.log_density([40.0], beta=0.25)
    assert d["noos"] == 1 and d["loglik"] == 0.0
    assert abs(d["logdens"] - (d["fixed"] - 0.25 * 1e100)) <= 1e85
    assert ml.log_density([40.0], beta=1.0)["logdens"] == -math.inf
    assert ml.log_density([40.0], beta=0.25, anneal_support=False)["logdens"] == -math.inf
    assert ml.log_density([40.0], beta=0.0)["logdens"] == ml.log_density([40.0], beta=0.0)["fixed"]


def test_cache_consistency_after_run(oracle, datasets):
    """SampledModel.check (SampledModel.java:178-188) holds during every PT test of the reference."""
    for spec in (datasets.normal_meanvar(n=200), datasets.logistic_regression(300, 4), datasets.gaussian_mixture(400),
                 datasets.ising(4), datasets.beta_binomial()):
        pt = oracle.PT(oracle.Model(spec), nChains=4, seed=5)
        pt.initialize()
        for _ in range(15):
            pt.scan()
        assert pt.check_caches() < 1e-10


def test_ising_edges_and_enumeration(oracle, datasets):
    """F/Functions.xtend:8-23 gives 2N^2 edges; T/TestDiscreteModels.xtend:17-30 uses N=2 (16 states)."""
    m = oracle.Model(datasets.ising(2))
    n_edges = sum(1 for f in range(m.n_factors) if m.factor_is_annealed(f))
    assert n_edges == 8 and m.n_factors == 12 and m.n_samplers == 4
    J = math.log(1 + math.sqrt(2)) / 2
    z = 0.0
    for st in range(16):
        s = [(st >> k) & 1 for k in range(4)]
        d = m.log_density(ints=s, beta=1.0)
        z += math.exp(d["logdens"])
        pm = [2 * v - 1 for v in s]
        # N=2: each neighbour pair appears twice (direct + wrap edge)
        expect = J * 2 * (pm[0] * pm[1] + pm[2] * pm[3] + pm[0] * pm[2] + pm[1] * pm[3])
        assert abs(d["loglik"] - expect) < 1e-12
        assert abs(d["fixed"] - 4 * math.log(0.5)) < 1e-12
    assert abs(math.log(z) - oracle.lib().orc_ising_exact_log_z(2, J, 0.0, 1.0)) < 1e-12


def test_ising_log_z_agreement(oracle, datasets):
    """T/TestEndToEnd.xtend:139-167: stepping stone vs thermodynamic integration agree within 0.1;
    here both are also checked against exhaustive enumeration (factories/Exact.java:82-92)."""
    spec = datasets.ising(4)
    h = spec["hyper"]
    exact = oracle.lib().orc_ising_exact_log_z(4, h["coupling"], h["moment"], 1.0)
    pt = oracle.PT(oracle.Model(spec), nChains=8, seed=11, nThreads=4)
    pt.initialize()
    res = pt.perform_inference(2000, 0.5, want_samples=False)
    st = pt.round_stats()
    assert abs(st["log_z_stepping_stone"] - exact) < 0.1
    assert abs(st["log_z_thermo"] - exact) < 0.1
    assert abs(st["log_z_stepping_stone"] - st["log_z_thermo"]) < 0.1
    assert res["rounds"][-1]["roundTrips"] > 0


def test_normal_posterior_against_quadrature(oracle, datasets):
    """C1 at reduced size: posterior moments and log Z against 2-D quadrature (independent truth)."""
    spec = datasets.normal_meanvar(n=100)
    y = spec["y"]; n = len(y)
    mu = np.linspace(y.mean() - 2.0, y.mean() + 2.0, 801)
    s2 = np.linspace(0.8, 14.0, 1601)
    MU, S2 = np.meshgrid(mu, s2, indexing="ij")
    ss = ((y[None, None, :] - MU[..., None]) ** 2).sum(-1)
    ll = -0.5 * n * np.log(2 * np.pi * S2) - 0.5 * ss / S2
    lp = -0.5 * np.log(2 * np.pi * 100.0) - 0.5 * MU ** 2 / 100.0 + np.log(0.1) - 0.1 * S2
    w = np.exp(ll + lp - (ll + lp).max())
    dA = (mu[1] - mu[0]) * (s2[1] - s2[0])
    logz = math.log(w.sum() * dA) + (ll + lp).max()
    e_mu, e_s2 = (w * MU).sum() / w.sum(), (w * S2).sum() / w.sum()
    pt = oracle.PT(oracle.Model(spec), nChains=8, seed=3, nThreads=4)
    pt.initialize()
    res = pt.perform_inference(4000, 0.5)
    smp = res["samples"][-2000:]
    st = pt.round_stats()
    assert abs(smp[:, 0].mean() - e_mu) < 0.05
    assert abs(smp[:, 1].mean() - e_s2) < 0.15
    assert abs(st["log_z_stepping_stone"] - logz) < 0.3
    # (thermodynamic integration is a trapezoid over 8 betas: too biased here to pin, cf. Ising test)


def test_adaptation_equalises_rejection(oracle):
    """EngineStaticUtils.java:124-168: the schedule generator maps equal steps of normalised cumulative
    Lambda to the new ladder; a ladder that is already equi-rejection is a fixed point."""
    betas = np.array([0.0, 0.1, 0.3, 0.6, 1.0])
    acc = np.array([0.7, 0.7, 0.7, 0.7])
    new, cum = oracle.adapt_ladder(betas, acc)
    assert np.allclose(new, betas, atol=1e-12) and np.allclose(cum, [0, .3, .6, .9, 1.2])
    acc = np.array([0.9, 0.5, 0.5, 0.9])   # more rejection in the middle -> points move to the middle
    new, _ = oracle.adapt_ladder(betas, acc)
    assert new[0] == 0.0 and new[-1] == 1.0 and np.all(np.diff(new) > 0)
    assert new[1] > betas[1] and new[3] < betas[3]
    # zero intensities trigger the 1e-6 nudge instead of a spline with repeated knots
    new, _ = oracle.adapt_ladder(betas, np.array([0.99999, 0.99999, 0.3, 0.99999]))
    assert np.all(np.isfinite(new)) and np.all(np.diff(new) > 0)


def test_round_trip_counter_matches_path_replay(oracle, datasets):
    """Paths.xtend:40-50,83-117 replayed in Python from the swap indicators of one round."""
    pt = oracle.PT(oracle.Model(datasets.normal_meanvar(n=30)), nChains=4, seed=9)
    pt.initialize()
    inds = []
    for _ in range(400):
        _, ind = pt.scan()
        inds.append(ind.copy())
    n = 4
    paths = [[c] for c in range(n)]
    for ind in inds:
        just = False
        for c in range(n):
            if just:
                just = False
                continue
            if ind[c] == 1:
                paths[c], paths[c + 1] = paths[c + 1], paths[c]
                paths[c].append(c); paths[c + 1].append(c + 1)
                just = True
            else:
                paths[c].append(c)
    count = 0
    for p in paths:
        new = False
        for s in p:
            if s == 0 and new:
                new = False; count += 1
            elif s == n - 1:
                new = True
    assert count == pt.round_stats()["round_trips"] and count > 0


def test_automatic_and_manual_annealing_agree(oracle, datasets):
    """T/TestEndToEnd.xtend:169-186 (testCustomAnnealer): PT on F/CustomAnnealRef.bl (x | mu ~ Normal(mu, 1), annealed
    automatically through ExponentiatedFactor) and on F/CustomAnnealTest.bl (the same density written as a LogPotential
    times an AnnealingParameter, SampledModel.sumOtherAnnealed :219-225) with the reference's defaults (8 chains, 1000
    scans, seed 1, SCM initialisation) must end with the same beta = 1 logDensity to 1e-10.  In the oracle the two
    paths share no code below the factor list: exponent * cached pre-annealing sum vs an uncached factor that reads
    the exponent itself."""
    finals = []
    for manual in (False, True):
        m = oracle.Model(datasets.custom_anneal_pair(manual))
        assert m.n_reals == 1 and m.n_samplers == 1
        pt = oracle.PT(m, nChains=8, seed=1)
        pt.initialize_scm()
        res = pt.perform_inference(1000, 0.5)
        d = pt.densities()
        finals.append((d["logdens"][0], res["indicators"].copy(), pt.get_state(0)[0][0], pt.get_ladder().copy()))
    (l1, i1, mu1, lad1), (l2, i2, mu2, lad2) = finals
    assert abs(l1 - l2) <= 1e-10                      # the reference's own assertion
    assert np.array_equal(i1, i2) and mu1 == mu2      # stronger: the whole run is the same
    assert np.allclose(lad1, lad2, rtol=0, atol=1e-15)
    # the automatic model reports the pre-annealing log-likelihood; the manual one has none (it is all "other annealed")
    a = oracle.Model(datasets.custom_anneal_pair(False)).log_density([0.3], beta=0.4)
    b = oracle.Model(datasets.custom_anneal_pair(True)).log_density([0.3], beta=0.4)
    assert abs(a["logdens"] - b["logdens"]) <= 1e-12 and b["loglik"] == 0.0 and a["loglik"] != 0.0


def test_lngamma_switch_moves_every_lgamma_based_density_by_less_than_1e9(oracle, datasets):
    """bayonet's SpecialFunctions.lnGamma is not in the tree; the oracle uses libm lgamma and carries the
    Pike & Hill series bayonet is believed to use as a switch.  Every catalogue density that calls lnGamma -- the
    beta-binomial family (nine calls per factor), the hierarchical rocket models (Beta / Binomial laws), the i.i.d. laws
    with lnGamma (Poisson, Binomial, NegativeBinomial, StudentT, Gamma, HalfStudentT, YuleSimon), Gamma / Beta priors
    and the Dirichlet constant of the mixtures -- moves by less than 1e-9 relative under the switch, on states drawn
    from the priors.  So WHICHEVER of the two bayonet implements, the 1e-9 parity bar of north_star holds for the other.
    What stays unpinned: bit parity with bayonet, and the arithmetic of its Random."""
    L = oracle.lib()
    rng = np.random.default_rng(11)
    specs = [datasets.beta_binomial(0), datasets.hierarchical_rockets(9), datasets.rockets_latent(6), datasets.gaussian_mixture(200),
             datasets.mixture_with_indicators(n=20, K=3)]
    specs += [datasets.iid_observations(d, 300) for d in ("poisson", "binomial", "negbinomial", "studentt", "gamma", "halfstudentt", "yulesimon")]
    worst = 0.0
    for spec in specs:
        m = oracle.Model(spec)
        pt = oracle.PT(m, nChains=6, seed=int(rng.integers(1, 1000)))
        pt.initialize()                       # prior draws: lnGamma arguments spread over (0, ~100)
        states = [pt.get_state(p) for p in range(6)]
        vals = []
        for which in (0, 1):
            L.orc_set_lngamma(which)
            try:
                vals.append([m.log_density(re, it, beta=0.7) for re, it in states])
            finally:
                L.orc_set_lngamma(0)
        for a, b in zip(*vals):
            for key in ("logdens", "loglik", "fixed"):
                if math.isfinite(a[key]) and math.isfinite(b[key]):
                    worst = max(worst, abs(a[key] - b[key]) / max(1.0, abs(a[key])))
            assert a["noos"] == b["noos"]
    assert 0.0 < worst < 1e-9, worst          # the switch does something, and stays inside the bar


def test_doomsday_evidence_exact_and_the_jvm_transcript(oracle, datasets):
    """The one number in the reference tree that a JVM run of the reference produced for this path: the Getting Started page
    prints the console output of `blang --model Doomsday --model.rate 1.0 --model.y 1.2 --model.z NA`
    (R/runtime/internals/doc/contents/GettingStarted.xtend:248-263): "Normalization constant estimate:
    -1.8657991502743467".  F/Doomsday.bl: z ~ Exponential(rate), y | z ~ ContinuousUniform(0.0, z); its evidence is
    closed form, Z = int_y^inf rate e^(-rate z) / z dz = rate E1(rate y), log Z = -1.84258 at (1.0, 1.2): the transcript's
    Monte Carlo estimate is 0.023 away from it.  The oracle's PT (the reference's algorithm: stepping-stone estimator over
    adapted ladders) and its SCM initialisation (the estimator of the engine that produced the transcript) must agree with
    the exact value within their own Monte Carlo error, and therefore with the JVM's number as closely as it agrees
    with the truth."""
    from scipy.special import exp1
    jvm = -1.8657991502743467
    exact = math.log(exp1(1.2))
    assert abs(exact - (-1.84258)) < 1e-5 and abs(jvm - exact) < 0.03
    spec = datasets.doomsday(1.2, 1.0)
    # the annealed density of the model, factor by factor: prior Exponential(1) on z, then -log(z) and the support y <= z
    d = oracle.Model(spec).log_density([2.0])
    assert abs(d["loglik"] - (-math.log(2.0))) < 1e-15 and abs(d["fixed"] - (-2.0)) < 1e-12 and d["noos"] == 0
    assert oracle.Model(spec).log_density([1.1])["noos"] == 1                       # z < y: out of support
    est = []
    for seed in range(12):
        pt = oracle.PT(oracle.Model(spec), nChains=8, seed=300 + seed)
        pt.initialize(oracle.INIT_FORWARD)
        res = pt.perform_inference(1000, 0.5)
        est.append(res["rounds"][-1]["logZ"])
    est = np.array(est)
    se = est.std(ddof=1) / math.sqrt(len(est))
    assert abs(est.mean() - exact) < 4 * se + 1e-3, (est.mean(), exact, se)
    assert abs(est.mean() - jvm) < abs(jvm - exact) + 4 * se + 1e-3
    scm = []
    for seed in range(40):
        pt = oracle.PT(oracle.Model(spec), nChains=2, seed=500 + seed)
        scm.append(pt.initialize_scm(nParticles=1000)["logNormEstimate"])
    scm = np.array(scm)
    # E[Z_hat] = Z (T/TestSMCUnbiasness): compare on the Z scale
    zs = np.exp(scm)
    assert abs(zs.mean() - math.exp(exact)) < 4 * zs.std(ddof=1) / math.sqrt(len(zs))
    assert abs(scm.mean() - jvm) < 0.05
    # the JVM's single run as one draw from the SCM estimator's distribution (40 seeds here: mean -1.849, sd 0.060)
    assert abs(jvm - scm.mean()) < 2.5 * scm.std(ddof=1)
