"""N-GPU sharded runs of five families against an unsharded run on rank 0: same seed => identical round statistics.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_families.py
(measured on 2 B200s in round 1: all five agree)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
from blangsdk_b200.distributed import ShardedEngine
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
specs = [D.linear_regression(5000, 4), D.iid_observations("studentt", 3000), D.hierarchical_rockets(9), D.gaussian_mixture(4000),
         D.logistic_regression(5000, 5)]
ok = True
for spec in specs:
    N = 8
    eng = Engine(nChains=N, random=11, device=lr, rank=rank, worldSize=world)
    eng.set_model(spec); eng.initialize("FORWARD")
    sh = ShardedEngine(eng); sh.prepare(); sh.run_scans(12); sh.synchronize()
    st = eng.round_stats()
    if rank == 0:
        ref = Engine(nChains=N, random=11, device=lr)
        ref.set_model(spec); ref.initialize("FORWARD"); ref.run_scans(12, want=())
        rs = ref.round_stats()
        same = np.allclose(st["swapAcceptMean"], rs["swapAcceptMean"], rtol=1e-9, atol=1e-12) and np.allclose(st["energyMean"], rs["energyMean"], rtol=1e-9) and np.array_equal(st["roundTrips"], rs["roundTrips"])
        print(spec["family"], "sharded == single:", same, "evals", eng.counters()["evals"], ref.counters()["evals"], flush=True)
        ok = ok and same
        ref.close()
    eng.close()
dist.barrier()
if rank == 0: print("ALL OK" if ok else "MISMATCH")
dist.destroy_process_group()
