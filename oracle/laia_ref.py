"""TEST INFRASTRUCTURE: loader for the reference's own Laia planner, python/hetu/laia/laia.pyx,
cythonized from where it lies under /root/reference into oracle/_ref/laia*.so (oracle/Makefile,
target `ref`).  Git-ignored, travels to the GPU box."""
import glob
import os
import sys

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mod = None


def available():
    return bool(glob.glob(os.path.join(_REF, "laia*.so")))


def module():
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref/laia*.so is not built: run `make -C oracle ref`")
        sys.path.insert(0, _REF)
        try:
            import laia  # noqa: the reference module name
        finally:
            sys.path.remove(_REF)
        _mod = laia
    return _mod


class _Queue(object):
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def run(sample_embs, epoch_num, mini_batch_size, batch_num, nrank, cache_size):
    """Run laia.pyx's laia_scheduler for every rank -> per batch (plans, dist): plans[w] ascending
    key list, dist[w] list of sample positions.  mini_batch_size * nrank must be >= 8 (the
    reference's prange chunk size is batch // 8) and the sample count a multiple of the global
    batch (its wrap-around copy writes through an unbound view)."""
    emb = np.ascontiguousarray(sample_embs, dtype=np.intc)
    per_rank = []
    for rank in range(nrank):
        q = _Queue()
        module().laia_scheduler(emb, epoch_num, mini_batch_size, batch_num, nrank, rank, cache_size, q)
        assert q.items[-1] == -1
        per_rank.append(q.items[:-1])
    nb = len(per_rank[0]) // 2
    out = []
    for b in range(nb):
        plans = [sorted(int(k) for k in per_rank[w][2 * b]) for w in range(nrank)]
        dist = [[int(p) for p in per_rank[w][2 * b + 1]] for w in range(nrank)]
        out.append((plans, dist))
    return out
