"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run).

Sharded engine (chain slots block-partitioned over the ranks, mailboxes connected with connect_peers) against an
unsharded engine on rank 0, same seed: the swap indicators must be identical, energies / samples / round statistics
equal to re-association error (bit-identical when the kernels use the same CTA shape), on every rank.  Also covers
SCM initialisation and performInference with adaptation on the sharded engine (PT.java:85-113, 300-329)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from blangsdk_b200 import datasets as D
from blangsdk_b200.distributed import connect_peers
from blangsdk_b200.engine import Engine

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, float), np.asarray(b, float)
    m = np.isfinite(a) & np.isfinite(b)
    return bool(np.array_equal(np.isnan(a), np.isnan(b)) and np.all(np.abs(a[m] - b[m]) <= tol * np.maximum(1.0, np.abs(b[m]))))


def same_everywhere(x):
    """True when every rank holds the same array (the table, the accumulators and the outputs are replicated)"""
    t = torch.from_numpy(np.ascontiguousarray(np.nan_to_num(np.asarray(x, float), nan=-7.0))).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


cases = [("normal", D.normal_meanvar(n=300), {}), ("logistic", D.logistic_regression(6000, 5), {}),
         ("logistic-pool", D.logistic_regression(6000, 5), {"BLANGPT_POOL": "1"}),
         ("mixture", D.gaussian_mixture(3000), {}), ("linreg", D.linear_regression(4000, 4), {}),
         ("ising", D.ising(8), {}), ("iid", D.iid_observations("studentt", 2000), {})]
N = 4 * world
for name, spec, env in cases:
    os.environ.update(env)
    for init in ("FORWARD", "SCM"):
        eng = Engine(nChains=N, random=11, device=lr, rank=rank, worldSize=world, ctasPerChain=0 if env else 1)
        eng.set_model(spec)
        connect_peers(eng)
        eng.initialize(init, nParticles=24) if init == "SCM" else eng.initialize(init)
        # the beta = 1 sample travels with the exchange for real-valued states; integer (Ising) samples do not
        want = ("energy", "swapIndicators") + (() if spec["family"] == "ising" else ("samples",))
        out = eng.run_scans(5, want=want)
        res = eng.perform_inference(30, 0.5, want=("energy", "swapIndicators"))
        st = eng.round_stats()
        dens = eng.log_density()
        replicated = all(same_everywhere(v) for v in (out["energy"], out["swapIndicators"], res["energy"], st["energyMean"], dens["loglik"]))
        if "samples" in out:
            replicated = replicated and same_everywhere(out["samples"])
        verdict = replicated
        if rank == 0:
            ref = Engine(nChains=N, random=11, device=lr, ctasPerChain=0 if env else 1)
            ref.set_model(spec)
            ref.initialize(init, nParticles=24) if init == "SCM" else ref.initialize(init)
            o2 = ref.run_scans(5, want=want)
            r2 = ref.perform_inference(30, 0.5, want=("energy", "swapIndicators"))
            s2 = ref.round_stats()
            same = (np.array_equal(out["swapIndicators"], o2["swapIndicators"]) and close(out["energy"], o2["energy"])
                    and np.array_equal(res["swapIndicators"], r2["swapIndicators"]) and close(res["energy"], r2["energy"])
                    and close(st["swapAcceptMean"], s2["swapAcceptMean"]) and close(st["energyMean"], s2["energyMean"])
                    and np.array_equal(st["roundTrips"], s2["roundTrips"]) and close(eng.get_ladder(), ref.get_ladder())
                    and close(dens["loglik"], ref.log_density()["loglik"]))
            if "samples" in out:
                same = same and close(out["samples"], o2["samples"])
            verdict = verdict and same
            print("%-14s %-7s sharded == single: %s   replicated on all ranks: %s" % (name, init, same, replicated), flush=True)
            ref.close()
        ok = ok and verdict
        eng.close()
    for k in env:
        os.environ.pop(k, None)
flag = torch.tensor([1.0 if ok else 0.0]).cuda()
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if flag.item() == 1.0 else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
