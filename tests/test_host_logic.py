"""CPU-side checks of the product: the C-ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and its pure-host pieces (rounds, sharding plan) agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from blangsdk_b200 import engine
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return engine, engine.load_library()


def test_library_exports_every_declared_symbol():
    engine, L = _lib()
    header = open(os.path.join(ROOT, "include", "blangpt.h")).read()
    declared = set(re.findall(r"\b(blangpt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), "libblangpt.so does not export %s" % name
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)


def test_default_config_matches_reference_defaults():
    engine, L = _lib()
    cfg = engine.Config()
    L.blangpt_default_config(ctypes.byref(cfg))
    # ParallelTempering.java:25-40, PT.java:47-82, Runner.xtend:69-73
    assert cfg.nChains == 8 and cfg.nPassesPerScan == 3.0 and cfg.random == 1
    assert cfg.usePriorSamples == 1 and cfg.reversible == 0 and cfg.annealSupport == 1
    assert cfg.treatNaNAsNegativeInfinity == 0 and cfg.statisticRecordedMaxChainIndex == 100


def test_rounds_matches_oracle(oracle):
    engine, _ = _lib()
    for n, f in ((1000, 0.5), (10, 0.0), (37, 0.3), (5000, 0.9), (1, 0.5), (2, 0.5)):
        assert engine.rounds(n, f) == oracle.rounds(n, f)
    with pytest.raises(engine.BlangPTError):
        engine.rounds(10, 1.0)


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse to run (north_star: no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    engine, _ = _lib()
    with pytest.raises(engine.BlangPTError) as ei:
        engine.Engine(nChains=4)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "blangsdk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "blang_oracle" not in src, f


def test_shard_plan():
    from blangsdk_b200.distributed import owner_of_slot, replica_shard, shard_plan
    assert shard_plan(64, 1) == [(0, 64)]
    assert shard_plan(1024, 8) == [(i * 128, 128) for i in range(8)]
    with pytest.raises(ValueError):
        shard_plan(10, 4)
    assert [owner_of_slot(s, 8, 2) for s in range(8)] == [0, 0, 0, 0, 1, 1, 1, 1]
    got = [replica_shard(4096, 8, r) for r in range(8)]
    assert got == [(r * 512, 512) for r in range(8)]
    got = [replica_shard(10, 4, r) for r in range(4)]
    assert got == [(0, 3), (3, 3), (6, 2), (8, 2)] and sum(n for _, n in got) == 10


def test_datasets_are_deterministic(datasets):
    a, b = datasets.logistic_regression(500, 8), datasets.logistic_regression(500, 8)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.all(a["x"][:, 0] == 1.0)
    assert datasets.normal_meanvar()["y"].shape == (1000,)
    bb = datasets.beta_binomial_batch(4)
    assert bb["y"].shape == (4, 32) and np.array_equal(bb["y"][2], datasets.beta_binomial(2)["y"])
