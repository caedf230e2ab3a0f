"""world_size-2 gloo test of the multi-GPU host logic: each rank owns a block of chain
slots, all-gathers its (loglik, noos, fixed, evals) records with the product's gather_table, and must end
up with the same replicated table as a single process -- so the replicated swap pass (same Philox key on
every rank) takes the same decisions.  The per-rank exploration is played by the oracle (a test double);
what is under test is the product's plan + gather code path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_chains, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blangsdk_b200 import datasets
    from blangsdk_b200.distributed import TABLE_COLS, gather_table, shard_plan
    from oracle import oracle as O
    first, n_local = shard_plan(n_chains, world)[rank]
    spec = datasets.normal_meanvar(n=64)
    pt = O.PT(O.Model(spec), nChains=n_chains, seed=4)
    pt.initialize()
    tables = []
    for scan in range(5):
        # every rank explores only ITS chains' records; the others are unknown until the gather
        before = pt.densities()
        e, ind = pt.scan()
        d = pt.densities()
        full = np.stack([d["loglik"], d["noos"].astype(float), d["fixed"], np.zeros(n_chains)], axis=1)
        send = torch.from_numpy(full[first:first + n_local].copy())
        table = torch.full((n_chains, TABLE_COLS), float("nan"), dtype=torch.float64)
        gather_table(send, table)
        assert torch.equal(table, torch.from_numpy(full)), "gathered table differs from the single-process table"
        tables.append(table.numpy().copy())
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.stack(tables))
    dist.destroy_process_group()


def test_two_rank_gather_replicates_the_table(tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 8, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy"); b = np.load(tmp_path / "rank1.npy")
    assert a.shape == (5, 8, 4) and np.array_equal(a, b) and np.all(np.isfinite(a))


def test_gather_single_process_is_a_copy():
    from blangsdk_b200.distributed import gather_table
    send = torch.arange(12, dtype=torch.float64).view(3, 4)
    table = torch.zeros(3, 4, dtype=torch.float64)
    assert torch.equal(gather_table(send, table), send)
