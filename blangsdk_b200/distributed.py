"""Multi-GPU NRPT: chain slots sharded over ranks, one process per GPU.

The reference parallelises over chains with a thread pool inside one JVM
(R/engines/ParallelTempering.java:68,79).  Here chain SLOTS are block-partitioned
over ranks and never move; a swap permutes temperature positions, so per scan the
only cross-GPU traffic is the per-chain record (loglik, nOutOfSupport, fixed,
evals) -- 32 bytes per chain -- all-gathered over NCCL/NVLink.  Every rank then runs
the same swap kernel on the replicated table with the same Philox key and reaches
the same decisions: no second message.  Independent PT replicas (config C5) shard
by replica with no communication at all (`replica_shard`).

torch.distributed is plumbing only: it moves the table; all arithmetic is in
libblangpt.so.  The same code runs under gloo with CPU tensors for the host-logic
tests (tests/test_sharding_gloo.py).
"""
import torch
import torch.distributed as dist

TABLE_COLS = 4  # loglik, noos, fixed, evals-this-scan


def shard_plan(n_chains, world_size):
    """[(first_slot, n_local)] per rank: contiguous blocks, as blangpt_create lays them out."""
    if n_chains % world_size != 0:
        raise ValueError("nChains must be a multiple of the world size")
    per = n_chains // world_size
    return [(r * per, per) for r in range(world_size)]


def owner_of_slot(slot, n_chains, world_size):
    return slot // (n_chains // world_size)


def replica_shard(n_replicas, world_size, rank):
    """(first_replica, n_local) for replicas-only sharding; remainders go to the low ranks."""
    base, rem = divmod(n_replicas, world_size)
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def gather_table(send, table, group=None):
    """All-gather the local [n_local, 4] segments into the replicated [n_chains, 4] table (in slot order)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_gather_into_tensor(table.view(-1), send.view(-1), group=group)
    else:
        table.view(-1).copy_(send.view(-1))
    return table


def connect_peers(engine, group=None):
    """Exchange the ranks' mailbox handles (64 bytes each) over torch.distributed and connect them: from here on the
    chains' records travel GPU to GPU inside the exploration kernels, and Engine.run_scans / perform_inference /
    initialize work on the sharded engine with no per-scan host code (include/blangpt.h, blangpt_peer_connect)."""
    world = dist.get_world_size(group)
    mine = engine.peer_handle()
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", engine.cfg.device)
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
        out = torch.empty(64 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, t, group=group)
        blob = bytes(out.cpu().numpy().tobytes())
    else:
        objs = [None] * world
        dist.all_gather_object(objs, mine, group=group)
        blob = b"".join(objs)
    engine.peer_connect([blob[64 * r:64 * (r + 1)] for r in range(world)])


class _DevPtr:
    def __init__(self, ptr, n, itemsize=8, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def tensor_from_device_ptr(ptr, n_doubles, device):
    """float64 torch view of library-owned device memory (no copy)."""
    return torch.as_tensor(_DevPtr(ptr, n_doubles), device=device)


class ShardedEngine:
    """Drives one rank's Engine through explore -> all-gather -> swap."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.device = torch.device("cuda", engine.cfg.device)
        L, h = engine.L, engine.h
        n_local = L.blangpt_local_chains(h, None)
        self.n_local = int(n_local)
        self.table = tensor_from_device_ptr(L.blangpt_table_ptr(h, 0), engine.N * TABLE_COLS, self.device)
        self.send = tensor_from_device_ptr(L.blangpt_table_ptr(h, 1), self.n_local * TABLE_COLS, self.device)
        self.set_stream(torch.cuda.current_stream(self.device))

    def set_stream(self, stream):
        self.stream = stream
        self.engine._check(self.engine.L.blangpt_set_stream(self.engine.h, stream.cuda_stream))

    def _gather(self):
        if self.engine.cfg.worldSize > 1:
            gather_table(self.send, self.table, self.group)

    def prepare(self):
        """After set_state / initialize: evaluate local chains, replicate the table, prime the energies."""
        e = self.engine
        e._check(e.L.blangpt_evaluate(e.h))
        self._gather()
        e._check(e.L.blangpt_prime(e.h))

    def explore(self):
        self.engine._check(self.engine.L.blangpt_explore(self.engine.h))

    def exchange(self):
        self._gather()

    def swap(self, want=()):
        """DEO swap pass; `want` selects this scan's outputs (as Engine.run_scans, one scan)"""
        e = self.engine
        if not want:
            e._check(e.L.blangpt_swap(e.h, None))
            return None
        so, bufs = e._outputs(1, tuple(w for w in want if not (w == "samples" and e.n_ints)))
        import ctypes
        e._check(e.L.blangpt_swap(e.h, ctypes.byref(so)))
        return {k: v[0] for k, v in bufs.items()}

    def scan(self, want=()):
        self.explore()
        self.exchange()
        return self.swap(want)

    def run_scans(self, n):
        for _ in range(n):
            self.scan()

    def synchronize(self):
        self.engine._check(self.engine.L.blangpt_synchronize(self.engine.h))
