"""-m gpu parity tests: the CUDA path (called through the C ABI of include/blangpt.h) against the CPU
oracle on identical seeded inputs.  Bar: annealed log-densities within 1e-9 relative (fp64), integer
work (swap indicators, evaluation counts, out-of-support counts, spins, round trips) bit-exact."""
import math

import numpy as np
import pytest

from gpu_util import assert_trajectory_parity, make_pair, rel, run_both

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _specs(D):
    return {
        "normal": D.normal_meanvar(n=777),
        "logistic": D.logistic_regression(1501, 6),
        "mixture": D.gaussian_mixture(1203),
        "betabinomial": D.beta_binomial(3),
        "ising": D.ising(8),
    }


def _random_state(rng, spec, n_reals, n_ints, kind):
    fam = spec["family"]
    if fam == "normal":
        st = [rng.normal(1.0, 2.0), rng.gamma(2.0, 2.0)]
        if kind == "oos":
            st[1] = -abs(st[1])
        return st, None
    if fam == "logistic":
        w = rng.normal(0, 1.0, n_reals)
        if kind == "oos":
            w *= 60.0          # |eta| >> 36.7 -> p rounds to exactly 1.0 -> log(1-p) = -inf on y = 0 rows
        return w, None
    if fam == "mixture":
        K = n_reals // 3
        pi = rng.dirichlet(np.ones(K))
        var = rng.gamma(2.0, 1.0, K)
        if kind == "oos":
            var[:] = -np.abs(var)      # every component invalid -> log(0) = -inf on every row
        return np.concatenate([rng.normal(0, 4, K), var, pi]), None
    if fam == "betabinomial":
        st = [rng.gamma(2.0, 1.0), rng.gamma(2.0, 2.0)]
        if kind == "oos":
            st[0] = -0.5
        return st, None
    return None, rng.integers(0, 2, n_ints)


@pytest.mark.parametrize("fam", ["normal", "logistic", "mixture", "betabinomial", "ising"])
@pytest.mark.parametrize("kind", ["regular", "oos"])
def test_log_density_parity(oracle, datasets, fam, kind):
    """blangpt_log_density == SampledModel.logDensity / preAnnealedLogLikelihood / nOutOfSupport on identical states."""
    if fam == "ising" and kind == "oos":
        pytest.skip("spins outside {0,1} are rejected at the boundary")
    from blangsdk_b200.engine import Engine
    spec = _specs(datasets)[fam]
    n = 6
    eng = Engine(nChains=n, random=1)
    eng.set_model(spec)
    m = oracle.Model(spec)
    rng = np.random.default_rng(42)
    states = []
    for p in range(n):
        re, it = _random_state(rng, spec, eng.n_reals, eng.n_ints, kind)
        eng.set_state(p, re, it)
        states.append((re, it))
    got = eng.log_density()
    ladder = eng.get_ladder()[0]
    for p in range(n):
        ref = m.log_density(states[p][0], states[p][1], beta=ladder[p])
        assert got["noos"][0, p] == ref["noos"]
        assert rel(got["loglik"][0, p], ref["loglik"]) < TOL
        assert rel(got["fixed"][0, p], ref["fixed"]) < TOL
        if math.isinf(ref["logdens"]):
            assert got["logdens"][0, p] == ref["logdens"]
        else:
            assert abs(got["logdens"][0, p] - ref["logdens"]) <= TOL * max(1.0, abs(ref["logdens"]))
    if kind == "oos" and fam != "ising":
        assert got["noos"].sum() > 0
    eng.close()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 129, 1025])
def test_ragged_sizes(oracle, datasets, n):
    """odd / tiny observation counts: tails of the 16-byte vector loads and empty thread ranges"""
    for spec in (datasets.normal_meanvar(n=n), datasets.logistic_regression(n, 3), datasets.gaussian_mixture(n)):
        for ctas in (1, 2, 4):
            eng, ref = make_pair(oracle, spec, 4, seed=n, ctas=ctas)
            d, r = eng.log_density(), ref.densities()
            assert rel(d["loglik"][0], r["loglik"]) < TOL and np.array_equal(d["noos"][0], r["noos"])
            assert_trajectory_parity(eng, ref, 3)
            eng.close()


@pytest.mark.parametrize("fam", ["normal", "logistic", "mixture", "betabinomial", "ising"])
def test_trajectory_parity(oracle, datasets, fam):
    """Same Philox layout on both sides: swap decisions, evaluation counts and states must agree draw for draw."""
    spec = _specs(datasets)[fam]
    eng, ref = make_pair(oracle, spec, 8, seed=17)
    assert_trajectory_parity(eng, ref, 12)
    eng.close()


@pytest.mark.parametrize("ctas", [1, 2, 4, 8])
def test_cluster_sizes(oracle, datasets, ctas):
    """thread-block clusters of 1/2/4/8 CTAs per chain (DSMEM reduction) all reproduce the oracle"""
    for spec in (datasets.logistic_regression(4099, 5), datasets.gaussian_mixture(2050), datasets.normal_meanvar(n=999)):
        eng, ref = make_pair(oracle, spec, 6, seed=23, ctas=ctas)
        assert eng.ctas_per_chain == ctas
        assert_trajectory_parity(eng, ref, 6)
        eng.close()


@pytest.mark.parametrize("threads,ctas", [(512, 1), (1024, 1), (1024, 2), ("l2ahead", 1), ("l2ahead", 2)])
def test_logistic_cta_shapes(oracle, datasets, monkeypatch, threads, ctas):
    """the C2 kernel's shapes (512- and 1024-thread CTAs; the variant with L2 prefetch hints that set_model picks when
    the per-chain caches outgrow the L2) walk the same trajectory as the oracle"""
    if threads == "l2ahead":
        monkeypatch.setenv("BLANGPT_L2AHEAD", "1")
    else:
        monkeypatch.setenv("BLANGPT_LOGISTIC_THREADS", str(threads))
    spec = datasets.logistic_regression(4099, 5)
    eng, ref = make_pair(oracle, spec, 6, seed=29, ctas=ctas)
    assert_trajectory_parity(eng, ref, 6)
    eng.close()


@pytest.mark.parametrize("opts", [dict(usePriorSamples=False), dict(reversible=True), dict(annealSupport=False),
                                  dict(nPassesPerScan=2.5), dict(nPassesPerScan=0.5)])
def test_engine_options(oracle, datasets, opts):
    """PT options of the reference (ParallelTempering.java:25-40, PT.java:47-82) keep parity, including a
    non-integer nPassesPerScan whose sampler permutation straddles scans (SampledModel.java:265-272)."""
    for spec in (datasets.normal_meanvar(n=200), datasets.logistic_regression(400, 5), datasets.gaussian_mixture(300)):
        eng, ref = make_pair(oracle, spec, 5, seed=31, **opts)
        assert_trajectory_parity(eng, ref, 10)
        eng.close()


def test_single_chain(oracle, datasets):
    """nChains = 1 (the MCMC engine alias, factories/MCMC.java): no swap pass, ladder [1.0]"""
    eng, ref = make_pair(oracle, datasets.normal_meanvar(n=100), 1, seed=2)
    assert list(eng.get_ladder()[0]) == [1.0]
    out = assert_trajectory_parity(eng, ref, 8)
    assert out["swapIndicators"].sum() == 0
    eng.close()


def test_out_of_support_start(oracle, datasets):
    """a chain started out of support uses the -1e100 slice height (RealSliceSampler.java:143-148) and recovers"""
    spec = datasets.normal_meanvar(n=120)
    eng, ref = make_pair(oracle, spec, 4, seed=5, init="COPIES")
    for p in range(4):
        eng.set_state(p, [0.5, -0.2], None)     # the initial window [x0-U, x0-U+1] can reach sigma2 > 0
        ref.set_state(p, [0.5, -0.2], None)
    eng.initialize("COPIES")
    d = eng.log_density()
    assert np.all(d["noos"][0] == 240) and np.all(np.isinf(d["fixed"][0]))
    assert_trajectory_parity(eng, ref, 12)
    assert np.all(eng.log_density()["noos"] == 0)
    eng.close()


def test_replicas(oracle, datasets):
    """nReplicas independent PT replicas in one handle == that many oracle runs keyed by replica id"""
    from blangsdk_b200.engine import Engine
    R, N = 5, 4
    batch = datasets.beta_binomial_batch(R)
    eng = Engine(nChains=N, nReplicas=R, random=9)
    eng.set_model(batch)
    eng.initialize("FORWARD")
    out = eng.run_scans(15)
    for r in range(R):
        ref = oracle.PT(oracle.Model(datasets.beta_binomial(r)), nChains=N, seed=9, replica=r)
        ref.initialize()
        for s in range(15):
            e, ind = ref.scan()
            assert np.array_equal(out["swapIndicators"][s, r], ind)
            assert rel(out["energy"][s, r], e) < TOL
    eng.close()


def test_rounds_adaptation_and_diagnostics(oracle, datasets):
    """PT.performInference end to end: adapted ladder, Lambda, log Z, round trips, acceptance
    (PT.java:85-171,387-505) against the oracle's restatement."""
    from blangsdk_b200.engine import PT
    for spec in (datasets.normal_meanvar(n=150), datasets.logistic_regression(300, 4)):
        pt = PT(nChains=6, nScans=200, adaptFraction=0.5, random=13, initialization="FORWARD")
        pt.setSampledModel(spec)
        res = pt.performInference()
        ref = oracle.PT(oracle.Model(spec), nChains=6, seed=13)
        ref.initialize()
        rres = ref.perform_inference(200, 0.5)
        assert [r["nScans"] for r in res["rounds"]] == [r["nScans"] for r in rres["rounds"]]
        assert np.array_equal(res["swapIndicators"][:, 0, :], rres["indicators"])
        assert rel(res["energy"][:, 0, :], rres["energies"]) < TOL
        assert rel(res["samples"][:, 0, :], rres["samples"]) < TOL
        assert rel(res["ladder"][0], ref.get_ladder()) < TOL
        for a, b in zip(res["rounds"], rres["rounds"]):
            assert int(a["roundTrips"][0]) == b["roundTrips"]
            assert rel(a["Lambda"][0], b["Lambda"]) < TOL and rel(a["logZ"][0], b["logZ"]) < TOL
            assert rel(a["minAccept"][0], b["minAccept"]) < TOL and rel(a["meanAccept"][0], b["meanAccept"]) < TOL
        st, rst = res["finalStats"], ref.round_stats()
        assert rel(st["swapAcceptMean"][0], rst["swap_accept_mean"]) < TOL
        assert rel(st["energyMean"][0], rst["energy_mean"]) < TOL
        assert rel(st["energyVariance"][0], rst["energy_var"]) < 1e-7
        assert rel(st["energyCovariance"][0], rst["energy_cov"]) < 1e-7
        assert rel(st["logZThermodynamic"][0], rst["log_z_thermo"]) < TOL
        assert np.array_equal(st["logSumN"][0], rst["log_sum_n"])
        pt.engine.close()


def test_ising_spins_bit_exact(oracle, datasets):
    """checkerboard sweep vs the oracle's ORC_ORDER_CHECKERBOARD restatement: integer states identical"""
    for N in (2, 4, 16):
        eng, ref = make_pair(oracle, datasets.ising(N), 6, seed=N)
        run_both(eng, ref, 7)
        for p in range(6):
            _, it = eng.get_state(p)
            _, ri = ref.get_state(p)
            assert np.array_equal(it, ri)
        eng.close()


def test_determinism(datasets):
    """R/validation/DeterminismTest.java:24-51: same seed => identical trace, bit for bit, for every family, cluster
    size and kernel variant (compute-sanitizer's racecheck is not available on the pool: an unordered shared-memory
    or DSMEM exchange would show up here as a run-to-run difference, as the one in the C2 commit once did)"""
    from blangsdk_b200.engine import Engine
    cases = [(datasets.logistic_regression(5000, 6), dict(nChains=8), 0), (datasets.logistic_regression(5000, 6), dict(nChains=8), 4),
             (datasets.gaussian_mixture(6000), dict(nChains=6), 0), (datasets.gaussian_mixture(6000), dict(nChains=6), 4),
             (datasets.normal_meanvar(n=3000), dict(nChains=8), 2), (datasets.ising(8), dict(nChains=6), 0),
             (datasets.beta_binomial_batch(16), dict(nChains=6, nReplicas=16), 0),
             (datasets.iid_observations("studentt", 4000), dict(nChains=8), 2),
             (datasets.hierarchical_rockets(20), dict(nChains=8, nReplicas=8), 0)]
    for spec, kw, ctas in cases:
        outs = []
        for _ in range(3):
            eng = Engine(random=77, ctasPerChain=ctas, **kw)
            eng.set_model(spec)
            eng.initialize("FORWARD")
            want = ("energy", "swapIndicators") + (("samples",) if spec["family"] != "ising" else ("intSamples",))
            outs.append(eng.run_scans(12, want=want))
            eng.close()
        for k in outs[0]:
            assert np.array_equal(outs[0][k], outs[1][k]) and np.array_equal(outs[0][k], outs[2][k]), (spec["family"], ctas, k)


def test_split_phase_equals_run_scans(datasets):
    """explore -> (gather) -> swap, the multi-GPU path, on one rank gives the run_scans trajectory"""
    import torch
    from blangsdk_b200.distributed import ShardedEngine
    from blangsdk_b200.engine import Engine
    spec = datasets.normal_meanvar(n=300)
    a = Engine(nChains=8, random=3); a.set_model(spec); a.initialize("FORWARD")
    want = ("energy", "logDensity", "nOutOfSupport", "swapIndicators", "samples")
    ref = a.run_scans(9, want=want)
    b = Engine(nChains=8, random=3); b.set_model(spec); b.initialize("FORWARD")
    sh = ShardedEngine(b)
    sh.prepare()
    for s in range(9):                       # per-scan outputs of the split-phase path == run_scans' batch outputs
        got = sh.scan(want if s % 2 == 0 else ())
        if got is not None:
            for k in want:
                assert np.array_equal(got[k], ref[k][s]), (k, s)
    sh.synchronize()
    for p in range(8):
        assert np.array_equal(a.get_state(p)[0], b.get_state(p)[0])
    assert a.counters()["evals"] == b.counters()["evals"]
    a.close(); b.close()


def test_two_rank_emulation_matches_single_rank(datasets):
    """Two handles (rank 0/1 of worldSize 2) on one GPU, their table segments exchanged by hand, reproduce the
    single-rank run bit for bit: results do not depend on the GPU count."""
    import torch
    from blangsdk_b200.distributed import ShardedEngine
    from blangsdk_b200.engine import Engine
    spec = datasets.logistic_regression(1200, 5)
    N = 8
    one = Engine(nChains=N, random=21, ctasPerChain=2); one.set_model(spec); one.initialize("FORWARD")
    one.run_scans(8, want=())
    ranks = []
    for r in range(2):
        e = Engine(nChains=N, random=21, rank=r, worldSize=2, ctasPerChain=2)
        e.set_model(spec); e.initialize("FORWARD")
        ranks.append(ShardedEngine(e))

    def exchange():
        torch.cuda.synchronize()
        for dst in ranks:
            for src_rank, src in enumerate(ranks):
                seg = src.send
                dst.table.view(N, 4)[src_rank * (N // 2):(src_rank + 1) * (N // 2)].copy_(seg.view(N // 2, 4))
        torch.cuda.synchronize()

    for s in ranks:
        s.engine._check(s.engine.L.blangpt_evaluate(s.engine.h))
    exchange()
    for s in ranks:
        s.engine._check(s.engine.L.blangpt_prime(s.engine.h))
    for _ in range(8):
        for s in ranks:
            s.explore()
        exchange()
        for s in ranks:
            s.swap()
    for s in ranks:
        s.synchronize()
    lad = one.get_ladder()[0]
    for p in range(N):
        want = one.get_state(p)[0]
        got = None
        for s in ranks:
            try:
                got = s.engine.get_state(p)[0]
            except Exception:
                continue
        assert got is not None and np.array_equal(want, got)
    for s in ranks:
        s.engine.close()
    one.close()


def test_error_paths(datasets):
    from blangsdk_b200.engine import BlangPTError, Engine, PT
    eng = Engine(nChains=4)
    with pytest.raises(BlangPTError):
        eng.run_scans(1)                       # no model yet
    with pytest.raises(BlangPTError):
        eng.set_model(dict(family="poisson", hyper={}))
    eng.set_model(datasets.normal_meanvar(n=20))
    with pytest.raises(BlangPTError):
        eng.set_ladder([0.9, 0.5, 0.2, 0.0])   # must start at exactly 1.0
    with pytest.raises(BlangPTError):
        eng.set_state(9, [0.0, 1.0])
    eng.set_state(0, [float("nan"), 1.0])
    with pytest.raises(BlangPTError) as ei:
        eng.log_density()
    assert "NaN" in str(ei.value)
    with pytest.raises(BlangPTError) as ei:      # scans from a NaN coordinate report the error too (and terminate)
        eng.run_scans(2)
        eng.counters()
    assert "NaN" in str(ei.value)
    eng.close()
    with pytest.raises(BlangPTError):
        PT(targetAccept=1.5)


def test_target_accept_resizes_the_ladder(datasets):
    """--engine.targetAccept (PT.java:156-163): after the last adaptive round the ladder has
    max(2, int(Lambda / (1 - target))) points, every chain restarts from the beta = 1 state, and the final
    round's swap acceptance sits near the target."""
    from blangsdk_b200.engine import PT
    pt = PT(nChains=12, nScans=600, adaptFraction=0.5, random=6, initialization="FORWARD", targetAccept=0.6)
    pt.setSampledModel(datasets.normal_meanvar(n=300))
    res = pt.performInference(want=("energy",))
    rs = res["rounds"]
    lam = rs[-2]["Lambda"]
    assert rs[-1]["nChains"] == max(2, int(lam / (1 - 0.6))) == pt.engine.N
    lad = res["ladder"][0]
    assert lad[0] == 1.0 and lad[-1] == 0.0 and np.all(np.diff(lad) < 0)
    acc = res["finalStats"]["swapAcceptMean"][0]
    assert abs(acc.mean() - 0.6) < 0.15
    pt.engine.close()


def test_logit_term_accuracy_and_support():
    """The lean fp64 row term of the C2 kernels: exact form within 1e-13 of a long-double evaluation on
    sg >= -12, the reference's naive arithmetic (incl. -inf once p rounds to 1.0) beyond."""
    from blangsdk_b200.engine import debug_logit_terms
    rng = np.random.default_rng(5)
    z = np.concatenate([rng.normal(0, 4, 20000), np.linspace(-45, 45, 9001), rng.normal(0, 30, 5000),
                        [0.0, -0.0, 1e-300, -1e-300, 700.0, -700.0, 36.7, 36.8, -36.8, 12.0, -12.0, 11.999999, -12.000001]])
    y = rng.integers(0, 2, z.size).astype(np.int32)
    got = debug_logit_terms(z, y)
    sg = np.where(y == 1, z, -z).astype(np.longdouble)
    exact = (np.minimum(sg, 0) - np.log1p(np.exp(-np.abs(sg)))).astype(np.float64)
    with np.errstate(divide="ignore", over="ignore"):
        pr = 1.0 / (1.0 + np.exp(-z))
        naive = np.log(np.where(y == 1, pr, 1.0 - pr))
    lean = sg >= -12.0
    assert np.max(np.abs(got[lean] - exact[lean])) < 1e-13
    assert np.max(np.abs(got[lean] - naive[lean])) < 2e-11          # the two forms agree where the lean one is used
    fin = ~lean & np.isfinite(naive)
    assert np.array_equal(np.isinf(got[~lean]), np.isinf(naive[~lean]))   # same out-of-support rows
    assert np.max(np.abs(got[fin] - naive[fin]) / np.abs(naive[fin])) < 1e-12
    assert np.isinf(got[~lean]).sum() > 0


@pytest.mark.parametrize("fam", ["betabinomial", "normal"])
def test_fused_replica_kernel(oracle, datasets, fam, monkeypatch):
    """many small replicas run in the fused kernel (one warp per chain, scan loop + swaps inside one launch)
    and must still reproduce the oracle replica by replica; the unfused path must give the same bits"""
    from blangsdk_b200.engine import Engine
    R, N, S = 12, 6, 25
    if fam == "betabinomial":
        spec = datasets.beta_binomial_batch(R)
        one = lambda r: datasets.beta_binomial(r)
    else:
        spec = datasets.normal_meanvar(n=333)
        one = lambda r: spec
    outs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("BLANGPT_FUSED", fused)
        eng = Engine(nChains=N, nReplicas=R, random=19)
        eng.set_model(spec)
        eng.initialize("FORWARD")
        launches0 = eng.counters()["launches"]
        outs[fused] = eng.run_scans(S, want=("energy", "swapIndicators", "samples", "logDensity"))
        outs[fused]["launches"] = eng.counters()["launches"] - launches0
        outs[fused]["evals"] = eng.counters()["evals"]
        outs[fused]["stats"] = eng.round_stats()
        eng.close()
    assert outs["1"]["launches"] == 1 and outs["0"]["launches"] == 2 * S
    for k in ("energy", "swapIndicators", "samples", "logDensity"):
        if fam == "betabinomial":     # same 32-lane reduction tree in both kernels: identical bits
            assert np.array_equal(outs["1"][k], outs["0"][k]), k
        else:                          # 32 lanes vs 128 threads: summation order differs
            assert rel(outs["1"][k], outs["0"][k]) < TOL, k
    assert np.array_equal(outs["1"]["swapIndicators"], outs["0"]["swapIndicators"])
    assert outs["1"]["evals"] == outs["0"]["evals"]
    assert np.array_equal(outs["1"]["stats"]["roundTrips"], outs["0"]["stats"]["roundTrips"])
    out = outs["1"]
    total = 0
    for r in range(R):
        ref = oracle.PT(oracle.Model(one(r)), nChains=N, seed=19, replica=r)
        ref.initialize()
        for s in range(S):
            e, ind = ref.scan()
            assert np.array_equal(out["swapIndicators"][s, r], ind)
            assert rel(out["energy"][s, r], e) < TOL
        total += ref.n_evals()
        assert rel(out["stats"]["swapAcceptMean"][r], ref.round_stats()["swap_accept_mean"]) < TOL
    assert total == out["evals"]


@pytest.mark.parametrize("fam", ["normal", "logistic", "mixture", "betabinomial", "ising"])
def test_scm_initialisation(oracle, datasets, fam):
    """PT's default initialisation (PT.java:300-329): annealed SMC with adaptive temperatures, stratified
    resampling and one sampler step per temperature, then one particle per ladder point.  Same temperatures,
    same resampling rounds, same chain states as the oracle's restatement; then the scans agree too."""
    from blangsdk_b200.engine import Engine
    spec = {"normal": datasets.normal_meanvar(n=150), "logistic": datasets.logistic_regression(400, 4),
            "mixture": datasets.gaussian_mixture(300), "betabinomial": datasets.beta_binomial(2),
            "ising": datasets.ising(6)}[fam]
    N = 5
    eng = Engine(nChains=N, random=41)
    eng.set_model(spec)
    eng.initialize("SCM", nParticles=40, threshold=0.9)
    ref = oracle.PT(oracle.Model(spec), nChains=N, seed=41,
                    sweepOrder=oracle.ORDER_CHECKERBOARD if fam == "ising" else oracle.ORDER_REFERENCE)
    rs = ref.initialize_scm(nParticles=40, threshold=0.9)
    gs = eng.scm_summary()
    assert gs["nTemperatures"] == rs["nTemperatures"] and gs["nResamplings"] == rs["nResamplings"]
    assert gs["nTemperatures"] > 3
    assert abs(gs["logNormEstimate"] - rs["logNormEstimate"]) <= 1e-8 * max(1.0, abs(rs["logNormEstimate"]))
    for p in range(N):
        re, it = eng.get_state(p)
        rr, ri = ref.get_state(p)
        assert rel(re, rr) < 1e-8 and np.array_equal(it, ri)
    d, r = eng.log_density(), ref.densities()
    assert rel(d["loglik"][0], r["loglik"]) < 1e-8
    # informed start: the cold chains are near their target, far better than a prior draw
    assert d["loglik"][0, 0] > d["loglik"][0, N - 1] or fam == "ising"
    out, e_ref, i_ref = run_both(eng, ref, 5)
    assert np.array_equal(out["swapIndicators"][:, 0, :], i_ref)
    assert rel(out["energy"][:, 0, :], e_ref) < 1e-8
    eng.close()
