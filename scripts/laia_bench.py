"""Planner throughput: herald_b200's C++ Laia planner against the reference's Cython planner
(oracle/_ref/laia*.so = python/hetu/laia/laia.pyx) on the same host, same input; plans compared.
usage: laia_bench.py [W] [mini_batch] [batches] [cache_size] [threads]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from herald_b200.laia import LaiaScheduler
from oracle import laia_ref

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mini = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 400000
threads = int(sys.argv[5]) if len(sys.argv) > 5 else (os.cpu_count() or 8)
T, V = 26, 33762577
rng = np.random.default_rng(1)
S = W * mini * nb
emb = ((rng.zipf(1.05, (S, T)) - 1) % V)
emb = (emb * 20866421 % V).astype(np.int32)
s = LaiaScheduler()
s.start(emb, S, T, 1, mini, nb - 1, W, 0, cap, threads)
t0 = time.perf_counter()
ours = []
while s.step():
    ours.append((s.plan_of(0).tolist(), s.dist_of(0).tolist()))
t_ours = time.perf_counter() - t0
print("herald_b200 planner: %d batches of %d samples x %d tables, %d workers, snapshot capacity %d, %d threads: "
      "%.1f ms per batch (%.2f M samples/s)" % (len(ours), W * mini, T, W, cap, threads,
                                                1e3 * t_ours / len(ours), len(ours) * W * mini / t_ours / 1e6))
if laia_ref.available():
    class Q(object):
        def __init__(self): self.items = []
        def put(self, x): self.items.append(x)
    q = Q()
    t0 = time.perf_counter()
    laia_ref.module().laia_scheduler(emb, 1, mini, nb - 1, W, 0, cap, q)
    t_ref = time.perf_counter() - t0
    nbat = (len(q.items) - 1) // 2
    print("reference laia.pyx (8 OpenMP threads scoring, W threads plan/update): %.1f ms per batch (%.2f M samples/s); "
          "speed-up %.1fx" % (1e3 * t_ref / nbat, nbat * W * mini / t_ref / 1e6, t_ref / t_ours))
    same = all(sorted(int(k) for k in q.items[2 * b]) == ours[b][0] and list(q.items[2 * b + 1]) == ours[b][1]
               for b in range(nbat))
    print("plans identical:", same)
