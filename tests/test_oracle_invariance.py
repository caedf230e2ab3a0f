"""Exact-invariance tests of the oracle's samplers, after R/validation/ExactInvarianceTest.java:71-102,221-247
(the reference's main correctness tool, T/TestSDKDistributions.xtend:12-18): a test function of the
parameters has the same law under (a) the prior and (b) prior -> simulate data -> K posterior sampler steps.
Two-sample KS, Bonferroni-corrected.  The CUDA engine reproduces the oracle draw for draw (-m gpu tests), so
these laws carry over.  Plus the exhaustive 16-state check of T/TestDiscreteModels.xtend:17-30 for Ising N=2."""
import math

import numpy as np
import pytest
from scipy import stats

N_REP = 600
K_SCANS = 6
ALPHA = 0.005


def _eit(oracle, make_spec, draw_prior, simulate, n_reals, stat_fns, seed0):
    rng = np.random.default_rng(seed0)
    prior_vals, post_vals = [], []
    for rep in range(N_REP):
        theta = draw_prior(rng)
        data = simulate(rng, theta)
        m = oracle.Model(make_spec(data))
        pt = oracle.PT(m, nChains=1, seed=1000 + rep)
        pt.set_state(0, theta)
        for _ in range(K_SCANS):
            pt.scan()
        post, _ = pt.get_state(0)
        prior_vals.append([f(np.asarray(theta)) for f in stat_fns])
        post_vals.append([f(post) for f in stat_fns])
        assert pt.check_caches() < 1e-9
    prior_vals, post_vals = np.array(prior_vals), np.array(post_vals)
    for j in range(len(stat_fns)):
        p = stats.ks_2samp(prior_vals[:, j], post_vals[:, j]).pvalue
        assert p > ALPHA / len(stat_fns), "statistic %d: KS p = %.2e" % (j, p)


def test_eit_normal(oracle):
    def prior(rng):
        return [rng.normal(0.0, 2.0), rng.exponential(1.0 / 0.7)]
    def sim(rng, th):
        return rng.normal(th[0], math.sqrt(th[1]), 4)
    spec = lambda y: dict(family="normal", hyper=dict(m0=0.0, v0=4.0, rate=0.7), y=y)
    _eit(oracle, spec, prior, sim, 2, [lambda t: t[0], lambda t: t[1], lambda t: t[0] * t[1]], 1)


def test_eit_logistic(oracle):
    x = np.column_stack([np.ones(6), np.linspace(-1.5, 1.5, 6)])
    def prior(rng):
        return rng.normal(0.0, math.sqrt(1.5), 2)
    def sim(rng, th):
        return (rng.random(6) < 1.0 / (1.0 + np.exp(-(x @ th)))).astype(np.int32)
    spec = lambda y: dict(family="logistic", hyper=dict(v0=1.5), x=x, y=y)
    _eit(oracle, spec, prior, sim, 2, [lambda t: t[0], lambda t: t[1], lambda t: t[0] ** 2 + t[1] ** 2], 2)


def test_eit_beta_binomial(oracle):
    nt = np.array([12, 7, 20], dtype=np.int32)
    def prior(rng):
        return [rng.exponential(1.0 / 0.8), rng.exponential(1.0 / 0.8)]
    def sim(rng, th):
        p = rng.beta(th[0], th[1], 3)
        p = np.clip(p, 1e-300, 1 - 1e-16)
        return rng.binomial(nt, p).astype(np.int32)
    spec = lambda y: dict(family="betabinomial", hyper=dict(rate=0.8), y=y, ntrials=nt)
    _eit(oracle, spec, prior, sim, 2, [lambda t: t[0], lambda t: t[1], lambda t: t[0] / (t[0] + t[1])], 3)


def test_eit_mixture(oracle):
    K = 2
    def prior(rng):
        return np.concatenate([rng.normal(0.0, 2.0, K), rng.gamma(2.0, 1.0 / 2.0, K), rng.dirichlet(np.ones(K))])
    def sim(rng, th):
        z = rng.random(5) < th[4]            # component 0 with probability pi_0
        comp = np.where(z, 0, 1)
        return rng.normal(th[comp], np.sqrt(th[2 + comp]))
    spec = lambda y: dict(family="mixture", hyper=dict(K=K, m0=0.0, v0=4.0, gshape=2.0, grate=2.0, conc=1.0), y=y)
    _eit(oracle, spec, prior, sim, 6,
         [lambda t: t[0], lambda t: t[2], lambda t: t[4], lambda t: t[0] * t[4] + t[1] * t[5]], 4)


def _eit_mixed(oracle, make_spec, draw_prior, simulate, stat_fns, seed0):
    """exact-invariance test for models with integer latents: theta = (reals, ints)"""
    rng = np.random.default_rng(seed0)
    prior_vals, post_vals = [], []
    for rep in range(N_REP):
        re, it = draw_prior(rng)
        m = oracle.Model(make_spec(simulate(rng, re, it)))
        pt = oracle.PT(m, nChains=1, seed=3000 + rep)
        pt.set_state(0, re, it)
        for _ in range(K_SCANS):
            pt.scan()
        pre, pit = pt.get_state(0)
        prior_vals.append([f(np.asarray(re), np.asarray(it)) for f in stat_fns])
        post_vals.append([f(pre, pit) for f in stat_fns])
        assert pt.check_caches() < 1e-9
    prior_vals, post_vals = np.array(prior_vals), np.array(post_vals)
    for j in range(len(stat_fns)):
        p = stats.ks_2samp(prior_vals[:, j], post_vals[:, j]).pvalue
        assert p > ALPHA / len(stat_fns), "statistic %d: KS p = %.2e" % (j, p)


@pytest.mark.parametrize("K", [2, 12])
def test_eit_mixture_indicators(oracle, K):
    """F/MixtureModel.bl with latent indicators, after T/TestSDKDistributions' `mix` instance: the CategoricalSampler
    (fixed window [0, K); with K = 12 > 10 the window test of IntSliceSampler.accept runs), the simplex sampler seeing
    the n Categorical factors, and the mean / variance samplers seeing every observation."""
    n = 4
    def prior(rng):
        pi = rng.dirichlet(np.ones(K))
        re = np.concatenate([rng.normal(0.0, 2.0, K), rng.gamma(2.0, 1.0 / 2.0, K), pi])
        return re, rng.choice(K, size=n, p=pi).astype(np.int32)
    def sim(rng, re, z):
        return rng.normal(re[z], np.sqrt(re[K + z]))
    spec = lambda y: dict(family="mixture_indicators", hyper=dict(K=K, m0=0.0, v0=4.0, gshape=2.0, grate=2.0, conc=1.0), y=y)
    _eit_mixed(oracle, spec, prior, sim,
               [lambda r, z: r[0], lambda r, z: r[K], lambda r, z: r[2 * K], lambda r, z: float(z[0]), lambda r, z: float(np.sum(z == 0)),
                lambda r, z: r[z[0]]], 11 + K)


def test_eit_changepoint(oracle):
    """a DiscreteUniform latent moved by UniformSampler (fixed window [0, n + 1), n + 1 > 10)"""
    n = 14
    def prior(rng):
        return rng.normal(0.0, 2.0, 2), np.array([rng.integers(0, n + 1)], dtype=np.int32)
    def sim(rng, re, c):
        return rng.normal(np.where(np.arange(n) < c[0], re[0], re[1]), 1.0)
    spec = lambda y: dict(family="changepoint", hyper=dict(m0=0.0, v0=4.0, s2=1.0), y=y)
    _eit_mixed(oracle, spec, prior, sim, [lambda r, c: r[0], lambda r, c: r[1], lambda r, c: float(c[0]), lambda r, c: r[0] * c[0]], 29)


@pytest.mark.parametrize("order", ["reference", "checkerboard"])
def test_ising_16_states(oracle, datasets, order):
    """Ising N=2 at beta = 1: state frequencies of a long single-chain run against exact enumeration."""
    spec = datasets.ising(2)
    m = oracle.Model(spec)
    logp = np.array([m.log_density(ints=[(s >> k) & 1 for k in range(4)])["logdens"] for s in range(16)])
    p = np.exp(logp - logp.max()); p /= p.sum()
    pt = oracle.PT(m, nChains=1, seed=5, nPassesPerScan=1.0,
                   sweepOrder=oracle.ORDER_CHECKERBOARD if order == "checkerboard" else oracle.ORDER_REFERENCE)
    counts = np.zeros(16)
    n, thin = 40000, 4
    for t in range(n * thin):
        pt.scan()
        if t % thin == 0:
            _, s = pt.get_state(0)
            counts[int(s[0] + 2 * s[1] + 4 * s[2] + 8 * s[3])] += 1
    # the two ordered states (all 0 / all 1, 42% each) exchange slowly: fold each state with its global
    # flip (equal probability by symmetry at moment = 0) for the tight check, loose check unfolded
    freq = counts / n
    folded = freq + freq[::-1]
    assert np.max(np.abs(folded - (p + p[::-1]))) < 0.01
    assert np.max(np.abs(freq - p)) < 0.06


# ---- family 6: one exact-invariance test per law, as T/TestSDKDistributions.xtend does for R/distributions ----
N_OBS = 10              # enough data for power: with 10 observations a 30 % error in a Weibull shape fails the test
_TRIALS = np.array([5.0, 9.0, 14.0, 3.0, 7.0, 11.0, 2.0, 20.0, 6.0, 8.0])


def _pos(rng):          # Gamma(2, 1) draw, the prior of every positive parameter below
    return rng.gamma(2.0, 1.0)


_IID_EIT = {
    # dist: (priors, prior draw, simulator of N_OBS observations)
    "poisson": ([("gamma", 2.0, 1.0)], lambda r: [_pos(r)], lambda r, t: r.poisson(t[0], N_OBS)),
    "binomial": ([("beta", 2.0, 3.0)], lambda r: [r.beta(2.0, 3.0)], lambda r, t: r.binomial(_TRIALS.astype(int), t[0])),
    "negbinomial": ([("gamma", 2.0, 1.0), ("uniform", 0.0, 1.0)], lambda r: [_pos(r), r.uniform(0.0, 1.0)],
                    lambda r, t: r.negative_binomial(t[0], 1.0 - t[1], N_OBS)),
    "studentt": ([("gamma", 2.0, 1.0), ("normal", 0.0, 4.0), ("gamma", 2.0, 1.0)],
                 lambda r: [_pos(r), r.normal(0.0, 2.0), _pos(r)], lambda r, t: t[1] + t[2] * r.standard_t(t[0], N_OBS)),
    "geometric": ([("beta", 2.0, 3.0)], lambda r: [r.beta(2.0, 3.0)], lambda r, t: r.geometric(t[0], N_OBS) - 1),
    "laplace": ([("normal", 0.0, 4.0), ("gamma", 2.0, 1.0)], lambda r: [r.normal(0.0, 2.0), _pos(r)],
                lambda r, t: r.laplace(t[0], t[1], N_OBS)),
    "exponential": ([("gamma", 2.0, 1.0)], lambda r: [_pos(r)], lambda r, t: r.exponential(1.0 / t[0], N_OBS)),
    "gamma": ([("gamma", 2.0, 1.0)] * 2, lambda r: [_pos(r), _pos(r)], lambda r, t: r.gamma(t[0], 1.0 / t[1], N_OBS)),
    "weibull": ([("gamma", 2.0, 1.0)] * 2, lambda r: [_pos(r), _pos(r)], lambda r, t: t[0] * r.weibull(t[1], N_OBS)),
    "gumbel": ([("normal", 0.0, 4.0), ("gamma", 2.0, 1.0)], lambda r: [r.normal(0.0, 2.0), _pos(r)],
               lambda r, t: r.gumbel(t[0], t[1], N_OBS)),
    "logisticdist": ([("normal", 0.0, 4.0), ("gamma", 2.0, 1.0)], lambda r: [r.normal(0.0, 2.0), _pos(r)],
                     lambda r, t: r.logistic(t[0], t[1], N_OBS)),
    "halfstudentt": ([("gamma", 2.0, 1.0)] * 2, lambda r: [_pos(r), _pos(r)], lambda r, t: np.abs(t[1] * r.standard_t(t[0], N_OBS))),
    "loglogistic": ([("gamma", 2.0, 1.0)] * 2, lambda r: [_pos(r), _pos(r)],
                    lambda r, t: t[0] * (lambda u: (u / (1.0 - u)) ** (1.0 / t[1]))(r.uniform(1e-12, 1 - 1e-12, N_OBS))),
    "gompertz": ([("gamma", 2.0, 1.0)] * 2, lambda r: [_pos(r), _pos(r)],
                 lambda r, t: t[1] * np.log(1.0 - np.log(1.0 - r.uniform(0.0, 1 - 1e-12, N_OBS)) / t[0])),
    "yulesimon": ([("gamma", 2.0, 1.0)], lambda r: [_pos(r)],
                  lambda r, t: r.geometric(np.exp(-r.exponential(1.0 / t[0], N_OBS))) - 1),
    "uniformmax": ([("gamma", 2.0, 1.0)], lambda r: [_pos(r)], lambda r, t: r.uniform(0.0, t[0], N_OBS)),
}


@pytest.mark.parametrize("dist", sorted(_IID_EIT))
def test_eit_iid(oracle, dist):
    priors, draw, sim = _IID_EIT[dist]
    trials = _TRIALS if dist == "binomial" else None
    spec = lambda y: dict(family="iid", hyper=dict(dist=dist, priors=priors), y=np.asarray(y, dtype=float), trials=trials)
    n_par = len(priors)
    fns = [(lambda t, j=j: t[j]) for j in range(n_par)] + [lambda t: float(np.prod(t[:n_par]))]
    _eit(oracle, spec, draw, sim, n_par, fns, 60 + sorted(_IID_EIT).index(dist))


def test_eit_hierarchical(oracle):
    """F/HierarchicalModel.bl un-marginalised: a, b and three group probabilities"""
    launches = np.array([6, 11, 17], dtype=np.int32)
    def prior(rng):
        a, b = rng.exponential(1.0), rng.exponential(1.0)
        p = np.clip(rng.beta(a, b, 3), 1e-300, 1 - 1e-16)
        return [a, b] + list(p)
    def sim(rng, th):
        return rng.binomial(launches, th[2:]).astype(np.int32)
    spec = lambda y: dict(family="hierarchical", hyper=dict(rate_a=1.0, rate_b=1.0), y=y, ntrials=launches)
    _eit(oracle, spec, prior, sim, 5, [lambda t: t[0], lambda t: t[1], lambda t: t[2], lambda t: t[0] / (t[0] + t[1])], 7)


def test_eit_linreg(oracle):
    """F/LinRegression.bl: slope, intercept and sigma"""
    x = np.column_stack([np.linspace(-2.0, 2.0, 8), np.ones(8)])
    def prior(rng):
        return list(rng.normal(0.0, math.sqrt(4.0), 2)) + [rng.uniform(0.0, 3.0)]
    def sim(rng, th):
        return x @ np.array(th[:2]) + th[2] * rng.standard_normal(8)
    spec = lambda y: dict(family="linreg", hyper=dict(v0=4.0, smax=3.0), x=x, y=y)
    _eit(oracle, spec, prior, sim, 3, [lambda t: t[0], lambda t: t[1], lambda t: t[2], lambda t: t[0] * t[2]], 8)


def test_eit_rockets_latent(oracle):
    """F/SimpleHierarchicalModel.bl: the latent counts move by IntSliceSampler with stepping out"""
    def prior(rng):
        a, b = rng.exponential(1.0), rng.exponential(1.0)
        p = np.clip(rng.beta(a + 0.5, b + 0.5, 3), 1e-300, 1 - 1e-16)
        return ([a, b] + list(p), list(rng.poisson(2.0, 3)))
    rng0 = np.random.default_rng(77)
    prior_vals, post_vals = [], []
    for rep in range(N_REP):
        th, L = prior(rng0)
        y = rng0.binomial(1 + np.array(L), th[2:]).astype(np.int32)
        m = oracle.Model(dict(family="rockets_latent", hyper=dict(rate_a=1.0, rate_b=1.0, pois_mean=2.0), y=y))
        pt = oracle.PT(m, nChains=1, seed=3000 + rep)
        pt.set_state(0, th, L)
        for _ in range(K_SCANS):
            pt.scan()
        re, it = pt.get_state(0)
        fns = lambda r, i: [r[0], r[1], r[2], float(i[0]), float(i[1] + i[2])]
        prior_vals.append(fns(th, L)); post_vals.append(fns(re, it))
        assert pt.check_caches() < 1e-9
    prior_vals, post_vals = np.array(prior_vals), np.array(post_vals)
    for j in range(prior_vals.shape[1]):
        p = stats.ks_2samp(prior_vals[:, j], post_vals[:, j]).pvalue
        assert p > ALPHA / prior_vals.shape[1], "statistic %d: KS p = %.2e" % (j, p)


def test_scm_normalisation_estimate_is_unbiased(oracle):
    """After T/TestSMCUnbiasness.xtend (exhaustive randomness on a two-state HMM there; Monte Carlo over seeds here): the
    annealed-SMC estimate of Z produced by the SCM initialisation (AdaptiveJarzynski.java:91-233, adaptive temperatures,
    stratified resampling) has expectation Z.  Model with a closed-form evidence: Poisson counts, Gamma(2, 1) prior."""
    rng = np.random.default_rng(5)
    y = rng.poisson(3.0, 12).astype(float)
    a0, b0, S, n = 2.0, 1.0, y.sum(), len(y)
    log_z = (a0 * math.log(b0) - math.lgamma(a0) + math.lgamma(a0 + S) - (a0 + S) * math.log(b0 + n)
             - sum(math.lgamma(v + 1) for v in y))
    spec = dict(family="iid", hyper=dict(dist="poisson", priors=[("gamma", a0, b0)]), y=y, trials=None)
    m = oracle.Model(spec)
    ratios = []
    for seed in range(400):
        pt = oracle.PT(m, nChains=2, seed=seed)
        rs = pt.initialize_scm(nParticles=30, threshold=0.8)
        ratios.append(math.exp(rs["logNormEstimate"] - log_z))
    ratios = np.array(ratios)
    se = ratios.std(ddof=1) / math.sqrt(len(ratios))
    assert abs(ratios.mean() - 1.0) < 4 * se + 1e-3, (ratios.mean(), se)
    assert se < 0.05                        # and it is a usable estimator, not just an unbiased one


def test_forward_generators_moments(oracle):
    """After T/TestMoments.java:166-171 (3 sigma / sqrt(n) on i.i.d. draws): the beta = 0 chain of a PT run is an exact
    prior draw every scan (SampledModel.forwardSample), so its states over many scans are i.i.d. from the priors; their
    first two moments must match the closed forms of Normal, Gamma, Exponential, Beta, Uniform, Dirichlet and Poisson."""
    n = 6000
    def prior_draws(spec, n_chains=2):
        pt = oracle.PT(oracle.Model(spec), nChains=n_chains, seed=17)
        pt.initialize(oracle.INIT_FORWARD)
        re, it = [], []
        for _ in range(n):
            pt.scan()
            r, i = pt.get_state(n_chains - 1)
            re.append(r.copy()); it.append(i.copy())
        return np.array(re), np.array(it)
    def check(x, mean, var, what):
        se_m = math.sqrt(var / len(x))
        assert abs(x.mean() - mean) < 4 * se_m, (what, x.mean(), mean)
        assert abs(x.var() - var) < 0.12 * var + 1e-12, (what, x.var(), var)
    y = np.array([1.0, 2.0])
    re, _ = prior_draws(dict(family="iid", hyper=dict(dist="studentt", priors=[("gamma", 2.5, 0.7), ("normal", -1.0, 4.0), ("exponential", 0.4)]), y=y, trials=None))
    check(re[:, 0], 2.5 / 0.7, 2.5 / 0.49, "gamma"); check(re[:, 1], -1.0, 4.0, "normal"); check(re[:, 2], 2.5, 6.25, "exponential")
    re, _ = prior_draws(dict(family="iid", hyper=dict(dist="negbinomial", priors=[("gamma", 0.6, 1.5), ("uniform", 0.2, 0.9)]), y=y, trials=None))
    check(re[:, 0], 0.4, 0.6 / 2.25, "gamma, shape < 1"); check(re[:, 1], 0.55, 0.49 / 12, "uniform")
    re, _ = prior_draws(dict(family="iid", hyper=dict(dist="geometric", priors=[("beta", 2.0, 3.0)]), y=y, trials=None))
    check(re[:, 0], 0.4, 6.0 / (25 * 6), "beta")
    re, _ = prior_draws(dict(family="mixture", hyper=dict(K=3, m0=0.5, v0=2.0, gshape=2.0, grate=2.0, conc=1.0), y=y))
    check(re[:, 0], 0.5, 2.0, "mixture mean"); check(re[:, 3], 1.0, 0.5, "mixture variance (gamma)")
    check(re[:, 6], 1.0 / 3, 2.0 / (9 * 4), "dirichlet(1,1,1) coordinate")
    assert np.allclose(re[:, 6:9].sum(axis=1), 1.0)
    re, it = prior_draws(dict(family="rockets_latent", hyper=dict(rate_a=1.0, rate_b=1.0, pois_mean=2.0), y=np.array([0, 1])))
    check(it[:, 0].astype(float), 2.0, 2.0, "poisson"); check(re[:, 0], 1.0, 1.0, "exponential(1)")
    assert np.all((re[:, 2:] > 0) & (re[:, 2:] < 1))


def test_eit_has_power(oracle):
    """After T/TestExactTest.xtend (the reference checks that its exact-invariance test rejects F/BadNormal.bl and
    BadRealSliceSampler): the tests above must be able to fail.  A simulator that is unfaithful to the law by 30 % in
    one parameter is rejected."""
    priors, draw, _ = _IID_EIT["weibull"]
    unfaithful = lambda r, t: t[0] * r.weibull(t[1] * 1.3, N_OBS)
    spec = lambda y: dict(family="iid", hyper=dict(dist="weibull", priors=priors), y=np.asarray(y, dtype=float), trials=None)
    with pytest.raises(AssertionError):
        _eit(oracle, spec, draw, unfaithful, 2, [lambda t: t[0], lambda t: t[1], lambda t: t[0] * t[1]], 99)
