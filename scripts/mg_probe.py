"""Diagnostics (2 GPUs, torchrun): time the owner sync of a lookup whose misses are all LOCAL rows
against one whose misses are all REMOTE rows (read over NVLink from the IPC-mapped peer shard)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import herald_b200 as hb
    from herald_b200 import ps
    from herald_b200.cstable import CacheSparseTable
    def exchange(b):
        obj = [b]; dist.broadcast_object_list(obj, src=0); return obj[0]
    ps.group_init(rank, world, local, exchange)
    comm = hb.worker_init(local)
    V, D, N = 33762577, 128, 52000
    table = comm.InitTensor(0, ps.kCacheTable, V, D, ps.Normal, 0.0, 0.01, 123)
    cst = CacheSparseTable(3376258, V, D, 0, "lru", 0)
    cst.cache.reserve(1 << 20)
    cst.perf_enabled(True)
    dev = hb.gpu(local)
    half = V // world + 1
    rng = np.random.default_rng(5 + rank)
    out = []
    for rep in range(6):
        for name, lo in (("local", rank * half), ("remote", ((rank + 1) % world) * half)):
            ids = np.unique(rng.integers(lo + 1000, lo + half - 2000, N)).astype(np.float32)
            k = hb.array(ids, dev); d = hb.empty((ids.size, D), dev)
            comm.BarrierWorker()
            if rank == 0 or os.environ.get("BOTH", "1") == "1":
                cst.embedding_lookup(k, d, sync=True)
                p = cst.perf[-1]
                out.append((name, rep, int(p["num_unique"]), int(p["num_transfered"]), round(1e3 * p["transfer_time"], 1),
                            round(1e3 * p["time"], 1)))
            comm.BarrierWorker()
    for o in out:
        print("rank", rank, *o, flush=True)
    del cst
    comm.ClearTensor(0)
    ps.group_finalize()
    dist.destroy_process_group()

main()
