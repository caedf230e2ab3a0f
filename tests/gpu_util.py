"""Helpers shared by the -m gpu parity tests: drive the CUDA engine (through the C ABI) and the oracle on
the same seeded inputs."""
import numpy as np


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    both_inf = (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))) | (np.isnan(a) & np.isnan(b))
    a, b = np.where(both_inf, 0.0, a), np.where(both_inf, 0.0, b)
    with np.errstate(invalid="ignore"):
        d = np.abs(np.where(both_inf, 0.0, a - b)) / np.maximum(1.0, np.abs(np.where(np.isinf(b), 1.0, b)))
    return float(np.max(d)) if d.size else 0.0


def make_pair(O, spec, nChains, seed, init="FORWARD", ctas=0, **opts):
    from blangsdk_b200.engine import Engine
    eng = Engine(nChains=nChains, random=seed, ctasPerChain=ctas, **opts)
    eng.set_model(spec)
    ref = O.PT(O.Model(spec), nChains=nChains, seed=seed,
               nPassesPerScan=opts.get("nPassesPerScan", 3.0), usePriorSamples=opts.get("usePriorSamples", True),
               reversible=opts.get("reversible", False), annealSupport=opts.get("annealSupport", True),
               sweepOrder=O.ORDER_CHECKERBOARD if spec["family"] == "ising" else O.ORDER_REFERENCE)
    if init == "FORWARD":
        eng.initialize("FORWARD")
        ref.initialize(O.INIT_FORWARD)
    return eng, ref


def run_both(eng, ref, scans):
    out = eng.run_scans(scans, want=("energy", "swapIndicators", "samples", "logDensity", "nOutOfSupport"))
    e_ref, i_ref = [], []
    for _ in range(scans):
        e, ind = ref.scan()
        e_ref.append(e); i_ref.append(ind)
    return out, np.array(e_ref), np.array(i_ref)


def assert_trajectory_parity(eng, ref, scans, tol=1e-9):
    out, e_ref, i_ref = run_both(eng, ref, scans)
    assert np.array_equal(out["swapIndicators"][:, 0, :], i_ref), "swap decisions differ"
    assert rel(out["energy"][:, 0, :], e_ref) < tol
    assert eng.counters()["evals"] == ref.n_evals(), "number of log-density evaluations differs"
    n = eng.N
    for p in range(n):
        re, it = eng.get_state(p)
        rr, ri = ref.get_state(p)
        assert rel(re, rr) < tol
        assert np.array_equal(it, ri)
    return out


def start_from(eng, ref, states):
    """the same explicit state in every chain of both engines (COPIES-style start from a given point)"""
    for p, st in enumerate(states):
        eng.set_state(p, st, None)
        ref.set_state(p, st, None)
    eng.initialize("COPIES")
