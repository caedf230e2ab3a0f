"""-m gpu: the persistent pool exploration kernel (blangsdk_b200/csrc/pool.cuh) against the oracle.

The pool kernel replaces one-cluster-per-chain exploration for the logistic family when there are enough chains;
BLANGPT_POOL=1 forces it at test sizes.  Same bar as every trajectory test: swap indicators identical to the
oracle's, equal numbers of log-density evaluations, energies and states within 1e-9 relative (fp64 tolerance of
north_star; sums are re-associated, nothing else differs)."""
import numpy as np
import pytest

from gpu_util import assert_trajectory_parity, make_pair, rel, start_from

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pool_on(monkeypatch):
    monkeypatch.setenv("BLANGPT_POOL", "1")


@pytest.mark.parametrize("n,p,chains", [(20_000, 8, 6), (4_097, 5, 3), (64, 3, 4), (1, 2, 2), (9_473, 32, 12), (130, 1, 1)])
def test_pool_trajectory_parity(oracle, datasets, pool_on, n, p, chains):
    spec = datasets.logistic_regression(n, p, seed=100 + n)
    eng, ref = make_pair(oracle, spec, chains, seed=11)
    assert eng.kernel_variant == 1
    assert_trajectory_parity(eng, ref, 6)
    eng.close()


@pytest.mark.parametrize("opts", [dict(nPassesPerScan=0.7), dict(nPassesPerScan=3.7, usePriorSamples=False),
                                  dict(reversible=True), dict(annealSupport=False)])
def test_pool_engine_options(oracle, datasets, pool_on, opts):
    spec = datasets.logistic_regression(6_000, 6, seed=77)
    eng, ref = make_pair(oracle, spec, 5, seed=5, **opts)
    assert eng.kernel_variant == 1
    assert_trajectory_parity(eng, ref, 5)
    eng.close()


def test_pool_copies_start_and_out_of_support(oracle, datasets, pool_on):
    """COPIES start from explicit states, one of them so far out that rows leave the support (-beta * 1e100 plateau)"""
    spec = datasets.logistic_regression(3_000, 4, seed=3)
    eng, ref = make_pair(oracle, spec, 4, seed=9, init=None)
    states = [np.zeros(4), np.array([0.3, -0.2, 0.1, 0.0]), np.array([60.0, -45.0, 80.0, 30.0]), np.full(4, -0.5)]
    start_from(eng, ref, states)
    assert_trajectory_parity(eng, ref, 4)
    eng.close()


def test_pool_matches_cluster_kernel(datasets, monkeypatch):
    """same seeds through both exploration kernels: identical decisions, energies within re-association error"""
    from blangsdk_b200.engine import Engine
    spec = datasets.logistic_regression(30_000, 10, seed=21)
    outs = []
    for variant in ("0", "1"):
        monkeypatch.setenv("BLANGPT_POOL", variant)
        eng = Engine(nChains=8, random=4)
        eng.set_model(spec)
        assert eng.kernel_variant == int(variant)
        eng.initialize("FORWARD")
        outs.append((eng.run_scans(8, want=("energy", "swapIndicators", "samples")), eng.counters()["evals"]))
        eng.close()
    (a, ea), (b, eb) = outs
    assert ea == eb
    assert np.array_equal(a["swapIndicators"], b["swapIndicators"])
    assert rel(a["energy"], b["energy"]) < 1e-10
    assert rel(a["samples"], b["samples"]) < 1e-9


def test_pool_determinism_and_many_chains(datasets, pool_on):
    """more chains than SMs (several chains per controller warp) and bit-identical repeat runs"""
    from blangsdk_b200.engine import Engine
    spec = datasets.logistic_regression(2_000, 3, seed=8)
    runs = []
    for _ in range(2):
        eng = Engine(nChains=200, random=3)
        eng.set_model(spec)
        assert eng.kernel_variant == 1
        eng.initialize("FORWARD")
        runs.append(eng.run_scans(3, want=("energy", "swapIndicators")))
        eng.close()
    assert np.array_equal(runs[0]["energy"], runs[1]["energy"])
    assert np.array_equal(runs[0]["swapIndicators"], runs[1]["swapIndicators"])


def test_pool_rounds_and_scm(oracle, datasets, pool_on):
    """performInference with adaptation (the round driver resets accumulators between pool launches) and SCM
    initialisation (particles run on the cluster kernel, chains on the pool)"""
    from blangsdk_b200.engine import PT
    spec = datasets.logistic_regression(2_500, 4, seed=31)
    pt = PT(nChains=6, nScans=60, adaptFraction=0.5, random=17, initialization="FORWARD")
    pt.setSampledModel(spec)
    assert pt.engine.kernel_variant == 1
    res = pt.performInference()
    ref = oracle.PT(oracle.Model(spec), nChains=6, seed=17)
    ref.initialize()
    rres = ref.perform_inference(60, 0.5)
    assert np.array_equal(res["swapIndicators"][:, 0, :], rres["indicators"])
    assert rel(res["energy"][:, 0, :], rres["energies"]) < 1e-9
    assert rel(res["ladder"][0], ref.get_ladder()) < 1e-9
    for a, b in zip(res["rounds"], rres["rounds"]):
        assert int(a["roundTrips"][0]) == b["roundTrips"]
        assert rel(a["Lambda"][0], b["Lambda"]) < 1e-9
    pt.engine.close()
    # SCM initialisation followed by pool scans
    eng, ref = make_pair(oracle, spec, 5, seed=23, init=None)
    eng.initialize("SCM", nParticles=20)
    ref.initialize_scm(nParticles=20)
    assert_trajectory_parity(eng, ref, 3)
    eng.close()


# ---- the mixture family (BASELINE config 2) on the pool kernel: blangsdk_b200/csrc/pool_mixture.cuh ----
@pytest.mark.parametrize("n,chains", [(300, 5), (4_097, 3), (1, 2), (130, 1), (20_000, 6), (63, 4)])
def test_pool_mixture_trajectory_parity(oracle, datasets, pool_on, n, chains):
    spec = datasets.gaussian_mixture(n, seed=40 + n)
    eng, ref = make_pair(oracle, spec, chains, seed=19)
    assert eng.kernel_variant == 1
    assert_trajectory_parity(eng, ref, 6)
    eng.close()


@pytest.mark.parametrize("opts", [dict(nPassesPerScan=0.7), dict(nPassesPerScan=3.7, usePriorSamples=False),
                                  dict(reversible=True), dict(annealSupport=False)])
def test_pool_mixture_engine_options(oracle, datasets, pool_on, opts):
    spec = datasets.gaussian_mixture(2_000, seed=5)
    eng, ref = make_pair(oracle, spec, 5, seed=8, **opts)
    assert eng.kernel_variant == 1
    assert_trajectory_parity(eng, ref, 5)
    eng.close()


def test_pool_mixture_out_of_range_rows_and_k8(oracle, datasets, pool_on):
    """states whose mixture sums underflow on many rows (the literal path, NaN marks in the scratch array), an invalid
    variance (out of support), and K = 8 components (the second template instance)"""
    spec = datasets.gaussian_mixture(1_500, seed=2)
    eng, ref = make_pair(oracle, spec, 4, seed=3, init=None)
    states = [[-6.0, -2.0, 2.0, 6.0, 1.0, 0.25, 0.25, 1.0, 0.25, 0.25, 0.25, 0.25],
              [300.0, 310.0, 320.0, 330.0, 1e-3, 1e-3, 1e-3, 1e-3, 0.1, 0.2, 0.3, 0.4],       # every component underflows
              [-6.0, -2.0, 2.0, 6.0, 1.0, -0.5, 0.25, 1.0, 0.25, 0.25, 0.25, 0.25],            # a negative variance
              [0.0, 0.0, 0.0, 0.0, 50.0, 60.0, 70.0, 80.0, 0.7, 0.1, 0.1, 0.1]]
    start_from(eng, ref, states)
    assert_trajectory_parity(eng, ref, 4)
    eng.close()
    spec8 = datasets.gaussian_mixture(900, K=8, seed=6)
    eng, ref = make_pair(oracle, spec8, 3, seed=4)
    assert eng.kernel_variant == 1
    assert_trajectory_parity(eng, ref, 4)
    eng.close()


def test_pool_mixture_matches_cluster_kernel_and_is_deterministic(datasets, monkeypatch):
    from blangsdk_b200.engine import Engine
    spec = datasets.gaussian_mixture(30_000, seed=21)
    outs = []
    for variant in ("0", "1", "1"):
        monkeypatch.setenv("BLANGPT_POOL", variant)
        eng = Engine(nChains=40, random=4)
        eng.set_model(spec)
        assert eng.kernel_variant == int(variant)
        eng.initialize("FORWARD")
        outs.append((eng.run_scans(4, want=("energy", "swapIndicators", "samples")), eng.counters()["evals"]))
        eng.close()
    (a, ea), (b, eb), (c, ec) = outs
    assert ea == eb == ec
    assert np.array_equal(a["swapIndicators"], b["swapIndicators"])
    assert rel(a["energy"], b["energy"]) < 1e-10 and rel(a["samples"], b["samples"]) < 1e-9
    assert np.array_equal(b["energy"], c["energy"]) and np.array_equal(b["samples"], c["samples"])      # bit-identical repeat


def test_pool_mixture_rounds_adaptation_against_the_oracle(oracle, datasets, pool_on):
    """a down-scaled twin of BASELINE config 2 (bimodal mixture, K = 4) through PT.performInference on the pool kernel:
    the adapted ladder, Lambda, log Z, swap acceptance and the round-trip count of every round equal the oracle's"""
    from blangsdk_b200.engine import PT
    spec = datasets.gaussian_mixture(1_500, seed=77)
    pt = PT(nChains=6, nScans=200, adaptFraction=0.5, random=31, initialization="FORWARD")
    pt.setSampledModel(spec)
    assert pt.engine.kernel_variant == 1
    res = pt.performInference()
    ref = oracle.PT(oracle.Model(spec), nChains=6, seed=31)
    ref.initialize()
    rres = ref.perform_inference(200, 0.5)
    assert [r["nScans"] for r in res["rounds"]] == [r["nScans"] for r in rres["rounds"]]
    assert np.array_equal(res["swapIndicators"][:, 0, :], rres["indicators"])
    assert rel(res["energy"][:, 0, :], rres["energies"]) < 1e-9
    assert rel(res["ladder"][0], ref.get_ladder()) < 1e-9
    for a, b in zip(res["rounds"], rres["rounds"]):
        assert int(a["roundTrips"][0]) == b["roundTrips"]
        assert rel(a["Lambda"][0], b["Lambda"]) < 1e-9 and rel(a["logZ"][0], b["logZ"]) < 1e-9
        assert rel(a["minAccept"][0], b["minAccept"]) < 1e-9 and rel(a["meanAccept"][0], b["meanAccept"]) < 1e-9
    pt.engine.close()
