"""Sharded (one process per GPU) throughput of a BASELINE config shape other than the bench's C2.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py C4
C4 = 2-D Ising 256 x 256, 1024 chains sharded over the N GPUs (BASELINE config 3); C3 = mixture, 256 chains."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blangsdk_b200 import datasets as D  # noqa: E402
from blangsdk_b200.distributed import ShardedEngine  # noqa: E402
from blangsdk_b200.engine import Engine  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "C4"
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if which == "C4":
        spec, n_chains, scans, warm, name = D.ising(256), 1024, 40, 5, "C4 ising 256x256, 1024 chains"
    else:
        spec, n_chains, scans, warm, name = D.gaussian_mixture(), 256, 3, 1, "C3 mixture K=4 1M obs, 256 chains"
    eng = Engine(nChains=n_chains, random=1, device=lr, rank=rank, worldSize=world)
    eng.set_model(spec)
    eng.initialize("FORWARD")
    sh = ShardedEngine(eng)
    sh.prepare()
    sh.run_scans(warm)
    sh.synchronize()
    c0 = eng.counters()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(sh.stream)
    sh.run_scans(scans)
    ev1.record(sh.stream)
    sh.synchronize()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    evals = eng.counters()["evals"] - c0["evals"]
    t = torch.tensor([ms, float(evals)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, evals = float(mx[0]), float(sm[1])
    st = eng.round_stats()
    if rank == 0:
        print(json.dumps(dict(config=name, n_gpus=world, chains_per_gpu=n_chains // world, scans=scans, ms_per_scan=ms / scans,
                              scans_per_s=1e3 * scans / ms, evals_per_s=evals / (ms * 1e-3),
                              round_trips=int(st["roundTrips"].sum()), Lambda=float(np.mean(st["Lambda"])))), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
