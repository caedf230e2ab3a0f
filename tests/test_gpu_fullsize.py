"""-m gpu tests at BASELINE.json's full sizes, where the oracle is too slow: size-independent properties
(annealing linearity, additivity over data splits, permutation invariance, symmetry, DEO parity pattern,
round trips replayed from the swap indicators) and an independent numpy evaluation of the likelihood."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _logit_loglik(x, y, w):
    s = np.where(y == 1, 1.0, -1.0) * (x @ w)
    return float(np.sum(np.minimum(s, 0.0) - np.log1p(np.exp(-np.abs(s)))))


def test_c2_full_size_against_numpy(datasets):
    """C2 (100k x 32, 64 chains): after real scans, every chain's reported log-likelihood equals a numpy
    evaluation at the reported state; annealed density = fixed + beta * loglik."""
    from blangsdk_b200.engine import Engine
    spec = datasets.logistic_regression()
    eng = Engine(nChains=64, random=3)
    eng.set_model(spec)
    eng.initialize("COPIES")
    out = eng.run_scans(3, want=("energy", "logDensity", "nOutOfSupport", "swapIndicators"))
    d = eng.log_density()                        # from-scratch evaluation of the current states
    lad = eng.get_ladder()[0]
    for p in range(0, 63, 7):                    # cold chains: no row anywhere near the literal (z < -12) path
        w, _ = eng.get_state(p)
        ref = _logit_loglik(spec["x"], spec["y"], w)
        assert abs(d["loglik"][0, p] - ref) <= TOL * abs(ref)
        assert d["noos"][0, p] == 0
        prior = float(np.sum(-0.5 * math.log(2 * math.pi * 10.0) - 0.5 * w ** 2 / 10.0))
        assert abs(d["fixed"][0, p] - prior) <= TOL * abs(prior)
        assert abs(d["logdens"][0, p] - (prior + lad[p] * ref)) <= TOL * abs(prior + lad[p] * ref)
    # DEO: scan t only attempts pairs i = (t mod 2) + 2k
    ind = out["swapIndicators"][:, 0, :]
    for t in range(ind.shape[0]):
        assert np.all(ind[t, (1 - t % 2)::2] == 0)
    eng.close()


@pytest.mark.parametrize("fam", ["normal", "mixture", "logistic"])
def test_additivity_and_permutation(datasets, fam):
    """loglik(A u B) = loglik(A) + loglik(B), and shuffling the observations changes nothing (to rounding)"""
    from blangsdk_b200.engine import Engine
    rng = np.random.default_rng(0)
    if fam == "normal":
        spec = datasets.normal_meanvar(n=200_001)
        split = lambda s, idx: dict(s, y=s["y"][idx])
        state = [1.3, 3.7]
    elif fam == "mixture":
        spec = datasets.gaussian_mixture(n=300_001)
        split = lambda s, idx: dict(s, y=s["y"][idx])
        state = [-5.5, -2.2, 1.7, 6.3, 1.1, 0.3, 0.2, 0.9, 0.2, 0.3, 0.1, 0.4]
    else:
        spec = datasets.logistic_regression(150_001, 16)
        split = lambda s, idx: dict(s, x=s["x"][idx], y=s["y"][idx])
        state = rng.normal(0, 0.3, 16)
    n = len(spec["y"])
    perm = rng.permutation(n)

    def loglik(sp):
        e = Engine(nChains=2, random=1)
        e.set_model(sp)
        for p in range(2):
            e.set_state(p, state)
        v = e.log_density()["loglik"][0]
        e.close()
        assert v[0] == v[1]
        return v[0]
    full = loglik(spec)
    a, b = loglik(split(spec, perm[: n // 3])), loglik(split(spec, perm[n // 3:]))
    shuffled = loglik(split(spec, perm))
    assert abs(full - (a + b)) <= 1e-11 * abs(full)
    assert abs(full - shuffled) <= 1e-11 * abs(full)


def test_c4_ising_symmetries(datasets):
    """256 x 256 lattice: all-equal configuration has energy J * 2N^2 exactly; global flip symmetry at moment 0"""
    from blangsdk_b200.engine import Engine
    spec = datasets.ising(256)
    J, N = spec["hyper"]["coupling"], 256
    eng = Engine(nChains=4, random=1)
    eng.set_model(spec)
    rng = np.random.default_rng(1)
    s = rng.integers(0, 2, N * N)
    eng.set_state(0, None, np.ones(N * N, dtype=np.int32))
    eng.set_state(1, None, np.zeros(N * N, dtype=np.int32))
    eng.set_state(2, None, s)
    eng.set_state(3, None, 1 - s)
    d = eng.log_density()
    assert d["loglik"][0, 0] == J * 2 * N * N and d["loglik"][0, 1] == J * 2 * N * N
    assert d["loglik"][0, 2] == d["loglik"][0, 3]
    g = s.reshape(N, N) * 2 - 1
    ref = J * float(np.sum(g * np.roll(g, -1, 0)) + np.sum(g * np.roll(g, -1, 1)))
    assert abs(d["loglik"][0, 2] - ref) <= 1e-12 * max(1.0, abs(ref))
    assert np.allclose(d["fixed"][0], N * N * math.log(0.5), rtol=1e-13)
    eng.close()


def test_c5_round_trips_replayed_from_indicators(datasets):
    """many replicas: the device-side index process (Paths.nRejuvenations) equals a host replay of the
    recorded swap indicators, replica by replica"""
    from blangsdk_b200.engine import Engine
    R, N, S = 64, 8, 300
    eng = Engine(nChains=N, nReplicas=R, random=4)
    eng.set_model(datasets.beta_binomial_batch(R))
    eng.initialize("FORWARD")
    out = eng.run_scans(S, want=("swapIndicators",))
    st = eng.round_stats()
    ind = out["swapIndicators"]
    for r in range(R):
        pos = list(range(N))                       # pos[particle]
        at = list(range(N))                        # particle at position
        flag = [p == N - 1 for p in range(N)]
        count = 0
        for t in range(S):
            assert np.all(ind[t, r, (1 - t % 2)::2] == 0)
            for i in range(N - 1):
                if ind[t, r, i]:
                    a, b = at[i], at[i + 1]
                    at[i], at[i + 1] = b, a
            for p in range(N):
                q = at[p]
                if p == 0 and flag[q]:
                    flag[q] = False; count += 1
                elif p == N - 1:
                    flag[q] = True
        assert count == st["roundTrips"][r]
    assert st["roundTrips"].sum() > 0
    eng.close()


def test_ising_gpu_statistics_against_enumeration(oracle, datasets):
    """Checkerboard NRPT on a 4x4 lattice: stepping-stone log Z within 0.1 of exhaustive enumeration
    (the reference's own bar, T/TestEndToEnd.xtend:139-167) and mean energy at beta=1 within MC error."""
    from blangsdk_b200.engine import PT
    spec = datasets.ising(4)
    h = spec["hyper"]
    exact = oracle.lib().orc_ising_exact_log_z(4, h["coupling"], h["moment"], 1.0)
    pt = PT(nChains=8, nScans=4000, adaptFraction=0.5, random=5, initialization="FORWARD")
    pt.setSampledModel(spec)
    res = pt.performInference(want=("energy",))
    st = res["finalStats"]
    assert abs(st["logZSteppingStone"][0] - exact) < 0.1
    assert abs(st["logZThermodynamic"][0] - exact) < 0.1
    # d/dbeta log Z(beta) at 1 = E[loglik]: finite-difference the exact enumeration
    eps = 1e-4
    dlogz = (oracle.lib().orc_ising_exact_log_z(4, h["coupling"], h["moment"], 1.0 + eps) -
             oracle.lib().orc_ising_exact_log_z(4, h["coupling"], h["moment"], 1.0 - eps)) / (2 * eps)
    e = -res["energy"][-2000:, 0, 0]              # loglik of the beta = 1 chain in the final round
    nb = 40
    bm = e[: len(e) // nb * nb].reshape(nb, -1).mean(1)
    se = bm.std(ddof=1) / math.sqrt(nb)
    assert abs(e.mean() - dlogz) < 5 * se + 0.05
    pt.engine.close()


def test_c1_reference_config_against_quadrature(datasets):
    """BASELINE config 0 exactly as the reference would run it (1 000 observations, 8 chains, nScans = 1000 => 1001
    scans, adaptFraction 0.5, nPassesPerScan 3, SCM initialisation, five seeds): posterior means and variances of
    (mu, sigma^2) and log Z against 2-D quadrature, gated by batch-means Monte Carlo errors."""
    import math
    from blangsdk_b200.engine import PT
    spec = datasets.normal_meanvar()
    y = spec["y"]; n = len(y)
    assert n == 1000
    mu = np.linspace(y.mean() - 0.5, y.mean() + 0.5, 1001)
    s2 = np.linspace(2.8, 5.6, 1401)
    ybar, S = y.mean(), ((y - y.mean()) ** 2).sum()
    MU, S2 = np.meshgrid(mu, s2, indexing="ij")
    ll = -0.5 * n * np.log(2 * np.pi * S2) - 0.5 * (S + n * (MU - ybar) ** 2) / S2
    lp = -0.5 * np.log(2 * np.pi * 100.0) - 0.5 * MU ** 2 / 100.0 + np.log(0.1) - 0.1 * S2
    w = np.exp(ll + lp - (ll + lp).max())
    dA = (mu[1] - mu[0]) * (s2[1] - s2[0])
    truth = dict(logz=math.log(w.sum() * dA) + (ll + lp).max(),
                 m_mu=(w * MU).sum() / w.sum(), m_s2=(w * S2).sum() / w.sum())
    truth["v_mu"] = (w * (MU - truth["m_mu"]) ** 2).sum() / w.sum()
    truth["v_s2"] = (w * (S2 - truth["m_s2"]) ** 2).sum() / w.sum()

    def batch_se(x):
        b = int(math.sqrt(len(x)))
        means = x[:b * (len(x) // b)].reshape(b, -1).mean(axis=1)
        return means.std(ddof=1) / math.sqrt(b)

    z = {k: [] for k in ("m_mu", "m_s2", "v_mu", "v_s2")}
    logzs = []
    for seed in range(1, 6):
        pt = PT(nChains=8, nScans=1000, adaptFraction=0.5, random=seed, initialization="SCM")
        pt.setSampledModel(spec)
        res = pt.performInference()
        assert sum(r["nScans"] for r in res["rounds"]) == 1001
        s = res["samples"][-500:, 0, :]              # the final, non-adaptive round
        for name, x, t in (("m_mu", s[:, 0], truth["m_mu"]), ("m_s2", s[:, 1], truth["m_s2"]),
                           ("v_mu", (s[:, 0] - s[:, 0].mean()) ** 2, truth["v_mu"]),
                           ("v_s2", (s[:, 1] - s[:, 1].mean()) ** 2, truth["v_s2"])):
            z[name].append((x.mean() - t) / batch_se(x))
        logzs.append(res["rounds"][-1]["logZ"][0])
        pt.engine.close()
    for name, zs in z.items():
        assert np.max(np.abs(zs)) < 5.0 and abs(np.mean(zs)) < 2.5, (name, zs)
    assert abs(np.mean(logzs) - truth["logz"]) < 4 * max(np.std(logzs, ddof=1) / math.sqrt(5), 0.02), (logzs, truth["logz"])


def test_c2_posterior_against_laplace(datasets):
    """The bench workload end to end (100k x 32, 64 chains): the beta = 1 chain's posterior mean and standard
    deviations against the Gaussian (Laplace) approximation at the MAP, which is accurate to O(1/n) here --
    MAP and Hessian by Newton's method in numpy, independent of the engine."""
    from blangsdk_b200.engine import Engine
    spec = datasets.logistic_regression()
    X, yv, v0 = spec["x"], spec["y"].astype(np.float64), spec["hyper"]["v0"]
    w = np.zeros(X.shape[1])
    for _ in range(30):                                   # Newton on the log posterior
        p = 1.0 / (1.0 + np.exp(-(X @ w)))
        g = X.T @ (yv - p) - w / v0
        H = (X * (p * (1 - p))[:, None]).T @ X + np.eye(len(w)) / v0
        step = np.linalg.solve(H, g)
        w = w + step
        if np.abs(step).max() < 1e-12:
            break
    sd = np.sqrt(np.diag(np.linalg.inv(H)))
    eng = Engine(nChains=64, random=3)
    eng.set_model(spec)
    eng.initialize("COPIES")                              # every chain from w = 0, as in bench.py
    eng.run_scans(120, want=())                           # burn-in
    out = eng.run_scans(260, want=("samples",))
    s = out["samples"][:, 0, :]
    # autocorrelated draws: gate the mean with batch means, the spread with a generous band
    b = 13
    means = s[:b * (len(s) // b)].reshape(b, -1, s.shape[1]).mean(axis=1)
    se = means.std(axis=0, ddof=1) / np.sqrt(b)
    assert np.all(np.abs(s.mean(axis=0) - w) < 5 * se + 0.1 * sd), np.max(np.abs(s.mean(axis=0) - w) / sd)
    assert np.all(s.std(axis=0) > 0.6 * sd) and np.all(s.std(axis=0) < 1.5 * sd)
    eng.close()


@pytest.mark.parametrize("pool", ["0", "1"])
def test_c2_full_size_against_the_oracle(oracle, datasets, monkeypatch, pool):
    """BASELINE config 1 at its full data size (100 000 x 32) against the ORACLE, draw for draw: four chains, one scan of
    96 sampler executions each (the oracle needs ~ 3 ms per evaluation, a few seconds in all), through both exploration
    kernels -- the cluster kernel and the pool kernel that bench.py times."""
    from gpu_util import assert_trajectory_parity, make_pair, start_from
    monkeypatch.setenv("BLANGPT_POOL", pool)
    spec = datasets.logistic_regression()
    eng, ref = make_pair(oracle, spec, 4, seed=17, init=None)
    assert eng.kernel_variant == int(pool)
    rng = np.random.default_rng(5)
    start_from(eng, ref, [rng.normal(0.0, 0.05, 32) for _ in range(4)])      # near the mode: the regime PT spends its time in
    assert_trajectory_parity(eng, ref, 1)
    eng.close()


@pytest.mark.parametrize("pool", ["0", "1"])
def test_c3_full_size_against_the_oracle(oracle, datasets, monkeypatch, pool):
    """BASELINE config 2 at its full data size (K = 4, 1 000 000 observations) against the ORACLE, draw for draw: two chains,
    two scans of three sampler executions each (nPassesPerScan = 0.34 of 9 samplers; ~ 50 ms per oracle evaluation).
    This is the size at which pass_first / pass_cached stream their per-chain scratch from HBM, two pairs per thread
    in flight, with the warp votes and the n_valid masks of the last, ragged batch."""
    from gpu_util import assert_trajectory_parity, make_pair, start_from
    monkeypatch.setenv("BLANGPT_POOL", pool)      # both exploration kernels: one CTA per chain, and the pool (pool_mixture.cuh)
    spec = datasets.gaussian_mixture()
    assert len(spec["y"]) == 1_000_000
    eng, ref = make_pair(oracle, spec, 2, seed=23, init=None, nPassesPerScan=0.34)
    assert eng.kernel_variant == int(pool)
    states = [[-6.1, -2.05, 1.9, 6.2, 1.1, 0.3, 0.2, 0.9, 0.25, 0.25, 0.25, 0.25],
              [-5.0, -1.0, 1.0, 5.0, 2.0, 1.0, 1.0, 2.0, 0.1, 0.4, 0.3, 0.2]]
    start_from(eng, ref, states)
    assert_trajectory_parity(eng, ref, 2)
    eng.close()
