"""-m gpu parity tests for SURVEY 8 row a9: CategoricalSampler and UniformSampler (IntSliceSampler with a fixed window,
R/mcmc/CategoricalSampler.xtend:24-28, UniformSampler.xtend:21-27, IntSliceSampler.java:53-132) on the device.
  family mixture_indicators  blang.examples.MixtureModel with latent cluster indicators; K = 12 > 10 runs the window
                             test IntSliceSampler.accept in fixed-window mode
  family changepoint         a DiscreteUniform latent (window n + 1 > 10)
The engine carries the integers as exact doubles after the reals; the oracle keeps them as ints."""
import numpy as np
import pytest

from gpu_util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


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


def _assert_parity(eng, ref, scans):
    nr = ref.model.n_reals
    out = eng.run_scans(scans, want=("energy", "swapIndicators", "samples"))
    for s in range(scans):
        e, ind = ref.scan()
        assert np.array_equal(out["swapIndicators"][s, 0], ind), "swap decisions differ at scan %d" % s
        assert rel(out["energy"][s, 0], e) < TOL
    assert eng.counters()["evals"] == ref.n_evals(), "number of log-density evaluations differs"
    for p in range(eng.N):
        st, _ = eng.get_state(p)
        rr, ri = ref.get_state(p)
        assert rel(st[:nr], rr) < TOL and np.array_equal(st[nr:].astype(np.int64), ri)


@pytest.mark.parametrize("K,n", [(2, 20), (3, 30), (12, 40), (16, 16)])
def test_mixture_indicators_trajectory_parity(oracle, datasets, K, n):
    spec = datasets.mixture_with_indicators(n=n, K=K)
    eng, ref = _pair(oracle, spec, 5, seed=13)
    assert eng.n_reals == 3 * K + n and eng.n_samplers == 2 * K + 1 + n == ref.model.n_samplers
    _assert_parity(eng, ref, 8)
    eng.close()


@pytest.mark.parametrize("n", [5, 14, 200, 4097, 40000])
def test_changepoint_trajectory_parity(oracle, datasets, n):
    spec = datasets.change_point(n=n)
    eng, ref = _pair(oracle, spec, 6, seed=21)
    _assert_parity(eng, ref, 6)
    eng.close()


@pytest.mark.parametrize("fam", ["mixture_indicators", "changepoint"])
def test_a9_options_and_log_density(oracle, datasets, fam):
    spec = datasets.mixture_with_indicators(n=24, K=11) if fam == "mixture_indicators" else datasets.change_point(n=300)
    for opts in (dict(usePriorSamples=False), dict(reversible=True, nPassesPerScan=0.7), dict(annealSupport=False)):
        eng, ref = _pair(oracle, spec, 4, seed=5, **opts)
        _assert_parity(eng, ref, 5)
        d, r = eng.log_density(), ref.densities()
        assert rel(d["loglik"][0], r["loglik"]) < TOL and rel(d["fixed"][0], r["fixed"]) < TOL
        assert np.array_equal(d["noos"][0], r["noos"])
        eng.close()


def test_a9_rounds_and_scm(oracle, datasets):
    from blangsdk_b200.engine import Engine
    for spec in (datasets.mixture_with_indicators(n=20, K=3), datasets.change_point(n=120)):
        eng = Engine(nChains=5, random=31)
        eng.set_model(spec)
        eng.initialize("SCM", nParticles=30)
        ref = oracle.PT(oracle.Model(spec), nChains=5, seed=31)
        rs = ref.initialize_scm(nParticles=30)
        gs = eng.scm_summary()
        assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
        _assert_parity(eng, ref, 4)
        eng.close()


def test_changepoint_recovers_the_change_point(datasets):
    """statistical sanity at beta = 1: the posterior of c concentrates near the generating change point (60 % of n)"""
    from blangsdk_b200.engine import Engine
    spec = datasets.change_point(n=400)
    eng = Engine(nChains=6, random=3)
    eng.set_model(spec)
    eng.initialize("FORWARD")
    eng.run_scans(200, want=())
    out = eng.run_scans(600, want=("samples",))
    c = out["samples"][:, 0, 2]
    assert abs(np.median(c) - 240) < 60, np.median(c)
    eng.close()


def test_automatic_and_manual_annealing_agree_on_the_device(oracle, datasets):
    """T/TestEndToEnd.xtend:169-186 on the engine: PT with the reference's defaults (8 chains, 1000 scans, seed 1, SCM
    initialisation) on F/CustomAnnealRef.bl and on F/CustomAnnealTest.bl ends with the same beta = 1 logDensity to 1e-10
    (here: the same run, bit for bit), and both walk the oracle's trajectories."""
    from blangsdk_b200.engine import PT
    finals = []
    for manual in (False, True):
        spec = datasets.custom_anneal_pair(manual)
        pt = PT(nChains=8, nScans=1000, adaptFraction=0.5, random=1, initialization="SCM")
        pt.setSampledModel(spec)
        res = pt.performInference(want=("swapIndicators", "samples", "logDensity"))
        ref = oracle.PT(oracle.Model(spec), nChains=8, seed=1)
        ref.initialize_scm()
        rres = ref.perform_inference(1000, 0.5)
        assert np.array_equal(res["swapIndicators"][:, 0, :], rres["indicators"])
        assert rel(res["samples"][:, 0, :], rres["samples"]) < TOL
        assert rel(res["ladder"][0], ref.get_ladder()) < TOL
        d = pt.engine.log_density()
        assert abs(d["logdens"][0, 0] - ref.densities()["logdens"][0]) <= 1e-9 * max(1.0, abs(d["logdens"][0, 0]))
        finals.append((d["logdens"][0, 0], res["swapIndicators"].copy(), res["samples"].copy()))
        pt.engine.close()
    assert abs(finals[0][0] - finals[1][0]) <= 1e-10
    assert np.array_equal(finals[0][1], finals[1][1]) and np.array_equal(finals[0][2], finals[1][2])
