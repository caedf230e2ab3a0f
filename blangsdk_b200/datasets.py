"""Synthetic data of the BASELINE.json config shapes.

Every generator is deterministic in its seed and returns a `spec` dict
{family, hyper, <arrays>} understood by blangsdk_b200.engine (and, in tests,
by the oracle).  Sizes default to the full configs; tests pass smaller ones.
"""
import math

import numpy as np

SEED_C1, SEED_C2, SEED_C3, SEED_C4, SEED_C5 = 20240001, 20240002, 20240003, 20240004, 20240005


def normal_meanvar(n=1000, seed=SEED_C1):
    """C1: y_i ~ N(1.5, 4); mu ~ Normal(0,100), sigma2 ~ Exponential(0.1)."""
    rng = np.random.default_rng(seed)
    y = rng.normal(1.5, 2.0, n)
    return dict(family="normal", hyper=dict(m0=0.0, v0=100.0, rate=0.1), y=y)


def logistic_regression(n=100_000, p=32, seed=SEED_C2, prior_var=10.0):
    """C2: X ~ N(0,1) with column 0 == 1, w* ~ N(0, 0.5^2), y ~ Bernoulli(logistic(X w*)); w_j ~ Normal(0, prior_var)
    (variance 10; the reference's own fixture F/SpikedGLM.bl:22 uses a Normal(0, 1) slab)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, p))
    x[:, 0] = 1.0
    w = rng.normal(0.0, 0.5, p)
    eta = x @ w
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.int32)
    return dict(family="logistic", hyper=dict(v0=float(prior_var)), x=x, y=y, w_true=w)


def gaussian_mixture(n=1_000_000, K=4, seed=SEED_C3):
    """C3: equal-weight N(-6,1), N(-2,.25), N(2,.25), N(6,1); indicators marginalised."""
    rng = np.random.default_rng(seed)
    means = np.array([-6.0, -2.0, 2.0, 6.0])
    sds = np.array([1.0, 0.5, 0.5, 1.0])
    if K != 4:
        means = np.linspace(-6.0, 6.0, K)
        sds = np.full(K, 0.75)
    comp = rng.integers(0, K, n)
    y = means[comp] + sds[comp] * rng.standard_normal(n)
    return dict(family="mixture", hyper=dict(K=K, m0=0.0, v0=100.0, gshape=1.0, grate=1.0, conc=1.0), y=y)


def ising(N=256, moment=0.0, coupling=None):
    """C4: F/Ising.bl on an N x N periodic lattice at the critical coupling log(1+sqrt 2)/2."""
    if coupling is None:
        coupling = math.log(1.0 + math.sqrt(2.0)) / 2.0
    return dict(family="ising", hyper=dict(N=N, coupling=coupling, moment=moment))


def beta_binomial(replica=0, groups=32, trials=50, seed=SEED_C5):
    """C5, one replica: theta_g ~ Beta(2,5), y_g ~ Binomial(50, theta_g); alpha, beta ~ Exponential(0.1)."""
    rng = np.random.default_rng(seed + replica)
    theta = rng.beta(2.0, 5.0, groups)
    y = rng.binomial(trials, theta).astype(np.int32)
    return dict(family="betabinomial", hyper=dict(rate=0.1), y=y, ntrials=np.full(groups, trials, dtype=np.int32))


def beta_binomial_batch(n_replicas=4096, groups=32, trials=50, seed=SEED_C5):
    """C5, all replicas stacked: y[r, g]."""
    ys = np.stack([beta_binomial(r, groups, trials, seed)["y"] for r in range(n_replicas)])
    return dict(family="betabinomial", hyper=dict(rate=0.1), y=ys,
                ntrials=np.full((n_replicas, groups), trials, dtype=np.int32))


SEED_IID = 20240006


def iid_observations(dist="poisson", n=2000, seed=SEED_IID):
    """Family "iid" (SURVEY 8(f)-4): y_i ~ Dist(theta) i.i.d. with one prior per parameter.  Parameters in the
    reference's declaration order: poisson (mean), binomial (p; trials are data), negbinomial (r, p),
    studentt (nu, mu, sigma), geometric (p), laplace (location, scale), exponential (rate), gamma (shape, rate),
    weibull (scale, shape), gumbel (location, scale), logisticdist (location, scale), halfstudentt (nu, sigma),
    loglogistic (scale, shape), gompertz (shape, scale), yulesimon (rho), uniformmax (max; ContinuousUniform(0, max))."""
    rng = np.random.default_rng(seed + sum(map(ord, dist)))
    spec = dict(family="iid", trials=None)
    if dist == "poisson":
        spec["y"] = rng.poisson(4.2, n).astype(np.float64)
        priors = [("exponential", 0.1)]
    elif dist == "binomial":
        tr = rng.integers(1, 40, n)
        spec["y"] = rng.binomial(tr, 0.3).astype(np.float64)
        spec["trials"] = tr.astype(np.float64)
        priors = [("beta", 2.0, 3.0)]
    elif dist == "negbinomial":
        spec["y"] = rng.negative_binomial(3.0, 0.4, n).astype(np.float64)   # numpy counts failures with success prob 0.4
        priors = [("gamma", 2.0, 0.5), ("uniform", 0.0, 1.0)]
    elif dist == "studentt":
        spec["y"] = 1.0 + 2.0 * rng.standard_t(4.0, n)
        priors = [("exponential", 0.1), ("normal", 0.0, 100.0), ("exponential", 0.5)]
    elif dist == "geometric":
        spec["y"] = (rng.geometric(0.25, n) - 1).astype(np.float64)          # failures before the first success
        priors = [("uniform", 0.0, 1.0)]
    elif dist == "laplace":
        spec["y"] = rng.laplace(-0.5, 1.5, n)
        priors = [("normal", 0.0, 25.0), ("gamma", 2.0, 1.0)]
    elif dist == "exponential":
        spec["y"] = rng.exponential(1.0 / 1.7, n)
        priors = [("gamma", 2.0, 1.0)]
    elif dist == "gamma":
        spec["y"] = rng.gamma(2.5, 1.0 / 1.2, n)
        priors = [("exponential", 0.5), ("exponential", 0.5)]
    elif dist == "weibull":
        spec["y"] = 2.0 * rng.weibull(1.5, n)
        priors = [("exponential", 0.3), ("exponential", 0.3)]
    elif dist == "gumbel":
        spec["y"] = rng.gumbel(0.5, 1.2, n)
        priors = [("normal", 0.0, 25.0), ("exponential", 0.5)]
    elif dist == "logisticdist":
        spec["y"] = rng.logistic(1.0, 0.8, n)
        priors = [("normal", 0.0, 25.0), ("exponential", 0.5)]
    elif dist == "halfstudentt":
        spec["y"] = np.abs(1.5 * rng.standard_t(4.0, n))
        priors = [("exponential", 0.1), ("exponential", 0.5)]
    elif dist == "loglogistic":
        u = rng.uniform(1e-6, 1 - 1e-6, n)
        spec["y"] = 2.0 * (u / (1.0 - u)) ** (1.0 / 3.0)
        priors = [("exponential", 0.3), ("exponential", 0.3)]
    elif dist == "gompertz":
        u = rng.uniform(0.0, 1 - 1e-9, n)
        spec["y"] = 1.5 * np.log(1.0 - np.log(1.0 - u) / 0.8)
        priors = [("exponential", 0.5), ("exponential", 0.5)]
    elif dist == "yulesimon":
        w = rng.exponential(1.0 / 2.5, n)
        spec["y"] = (rng.geometric(np.exp(-w)) - 1).astype(np.float64)
        priors = [("exponential", 0.3)]
    elif dist == "uniformmax":
        spec["y"] = rng.uniform(0.0, 3.0, n)
        priors = [("exponential", 0.4)]
    else:
        raise ValueError(dist)
    spec["hyper"] = dict(dist=dist, priors=priors)
    # the generating parameters, in declaration order (a well-conditioned place to start chains from)
    spec["truth"] = dict(poisson=[4.2], binomial=[0.3], negbinomial=[3.0, 0.6], studentt=[4.0, 1.0, 2.0], geometric=[0.25],
                         laplace=[-0.5, 1.5], exponential=[1.7], gamma=[2.5, 1.2], weibull=[2.0, 1.5], gumbel=[0.5, 1.2],
                         logisticdist=[1.0, 0.8], halfstudentt=[4.0, 1.5], loglogistic=[2.0, 3.0], gompertz=[0.8, 1.5],
                         yulesimon=[2.5], uniformmax=[3.2])[dist]
    return spec


def doomsday(y=1.2, rate=1.0):
    """F/Doomsday.bl, the model of the SDK's Getting Started page and of T/TestEndToEnd: z ~ Exponential(rate),
    y | z ~ ContinuousUniform(0.0, z), run there with --model.rate 1.0 --model.y 1.2 --model.z NA.  Its evidence is
    closed form: Z = rate * E1(rate * y)."""
    return dict(family="iid", trials=None, y=np.array([float(y)]), hyper=dict(dist="uniformmax", priors=[("exponential", float(rate))]),
                truth=[2.0 * y])


def hierarchical_rockets(groups=12, seed=20240007):
    """Family "hierarchical": F/HierarchicalModel.bl (the rocket-launch example), not marginalised."""
    rng = np.random.default_rng(seed)
    launches = rng.integers(1, 60, groups).astype(np.int32)
    p = rng.beta(1.2, 6.0, groups)
    failures = rng.binomial(launches, p).astype(np.int32)
    return dict(family="hierarchical", hyper=dict(rate_a=1.0, rate_b=1.0), y=failures, ntrials=launches)


def linear_regression(n=5000, p=3, seed=20240008):
    """Family "linreg": F/LinRegression.bl with p covariates (last column == 1, the intercept)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, p))
    x[:, p - 1] = 1.0
    w = rng.normal(0.0, 2.0, p)
    y = x @ w + 0.7 * rng.standard_normal(n)
    return dict(family="linreg", hyper=dict(v0=25.0, smax=10.0), x=x, y=y, truth=list(w) + [0.7])


def rockets_latent(groups=8, seed=20240009):
    """Family "rockets_latent": F/SimpleHierarchicalModel.bl -- launch counts are latent (Poisson(2)), only failures
    are observed.  State = a, b, the group probabilities, then the latent counts (as exact doubles)."""
    rng = np.random.default_rng(seed)
    launches = 1 + rng.poisson(2.0, groups)
    p = rng.beta(1.5, 3.0, groups)
    failures = rng.binomial(launches, p).astype(np.int32)
    return dict(family="rockets_latent", hyper=dict(rate_a=1.0, rate_b=1.0, pois_mean=2.0), y=failures)


def mixture_with_indicators(n=30, K=3, seed=20240010):
    """blang.examples.MixtureModel (F/MixtureModel.bl) with K components and latent cluster indicators (family
    `mixture_indicators`, SURVEY 8 row a9: CategoricalSampler; K > 10 runs IntSliceSampler's window test)."""
    rng = np.random.default_rng(seed)
    centres = np.linspace(-2.0 * (K - 1), 2.0 * (K - 1), K)
    z = rng.integers(0, K, n)
    y = rng.normal(centres[z], 0.7)
    return dict(family="mixture_indicators", y=y, hyper=dict(K=K, m0=0.0, v0=25.0, gshape=2.0, grate=2.0, conc=1.0))


def change_point(n=200, seed=20240011):
    """y_i ~ Normal(i < c ? mu_0 : mu_1, s2) with a latent change point c ~ DiscreteUniform(0, n + 1) (family
    `changepoint`, SURVEY 8 row a9: UniformSampler)."""
    rng = np.random.default_rng(seed)
    c = int(0.6 * n)
    y = np.concatenate([rng.normal(-0.4, 1.0, c), rng.normal(0.5, 1.0, n - c)])
    return dict(family="changepoint", y=y, hyper=dict(m0=0.0, v0=4.0, s2=1.0))


def custom_anneal_pair(manual, x=10.0):
    """F/CustomAnnealRef.bl (manual = False: x | mu ~ Normal(mu, 1) observed, annealed automatically) and
    F/CustomAnnealTest.bl (manual = True: the same density as a LogPotential multiplied by an AnnealingParameter):
    the pair of T/TestEndToEnd.xtend:169-186."""
    return dict(family="anneal_pair", x=float(x), hyper=dict(m0=0.0, v0=1.0, s2=1.0, manual=bool(manual)))
