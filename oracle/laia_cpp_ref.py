"""TEST INFRASTRUCTURE: loader for the reference's C++ planners — laia/src/laia_scheduler.cc and
laia/src/topk_scheduler.cc compiled unmodified from /root/reference into
oracle/_ref/laia_cache*.so (oracle/Makefile target `ref`; Boost replaced by the stand-ins under
oracle/ref_shim/boost).  Git-ignored, travels to the GPU box."""
import glob
import os
import sys

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mod = None


def available():
    return bool(glob.glob(os.path.join(_REF, "laia_cache*.so")))


def module():
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref/laia_cache*.so is not built: run `make -C oracle ref`")
        sys.path.insert(0, _REF)
        try:
            import laia_cache  # noqa: the reference module name
        finally:
            sys.path.remove(_REF)
        _mod = laia_cache
    return _mod


def _drain(sched):
    msgs = []
    while True:
        m = list(sched.pop())
        if m == [0]:
            # the reference's wire ends with {0} (topk_scheduler.cc:333); a PLAN that is exactly [0]
            # cannot be told from it — callers keep key 0 out of single-key plans
            break
        msgs.append([int(x) for x in m])
    assert len(msgs) % 2 == 0
    return [(msgs[2 * b], msgs[2 * b + 1]) for b in range(len(msgs) // 2)]


def _silenced(fn):
    """The reference planners print progress on stdout/stderr from their threads."""
    sys.stdout.flush()
    sys.stderr.flush()
    saved = os.dup(1), os.dup(2)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    os.dup2(devnull, 2)
    try:
        return fn()
    finally:
        os.dup2(saved[0], 1)
        os.dup2(saved[1], 2)
        for fd in saved + (devnull,):
            os.close(fd)


def run_topk(sample_embs, epoch_num, mini_batch_size, batch_num, nrank, cache_size, num_threads,
             dataset, top_k_table):
    """TopkScheduler (standalone mode) for every rank -> per batch (plans[w], dist[w])."""
    emb = np.ascontiguousarray(sample_embs, dtype=np.uint64)

    def one(rank):
        s = module().TopkScheduler()
        s.start(emb, emb.shape[0], emb.shape[1], epoch_num, mini_batch_size, batch_num, nrank, rank,
                cache_size, num_threads, dataset, top_k_table, False, 0, 1)
        out = _drain(s)
        del s
        return out

    per_rank = _silenced(lambda: [one(r) for r in range(nrank)])
    nb = len(per_rank[0])
    return [([per_rank[w][b][0] for w in range(nrank)], [per_rank[w][b][1] for w in range(nrank)])
            for b in range(nb)]


def run_laia(sample_embs, epoch_num, mini_batch_size, batch_num, nrank, cache_size, num_threads=4):
    """LaiaScheduler (the C++ one, laia/src/laia_scheduler.cc) for every rank."""
    emb = np.ascontiguousarray(sample_embs, dtype=np.uint64)

    def one(rank):
        s = module().LaiaScheduler()
        s.start(emb, emb.shape[0], emb.shape[1], epoch_num, mini_batch_size, batch_num, nrank, rank,
                cache_size, num_threads, 24)
        out = _drain(s)
        del s
        return out

    per_rank = _silenced(lambda: [one(r) for r in range(nrank)])
    nb = len(per_rank[0])
    return [([per_rank[w][b][0] for w in range(nrank)], [per_rank[w][b][1] for w in range(nrank)])
            for b in range(nb)]
