"""Generate tests/golden/laia_cases.npz from the REFERENCE's own planner.

Run in the authoring container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_laia.py
oracle/_ref/laia*.so is python/hetu/laia/laia.pyx (+ MiniLRUCache.h) cythonized unmodified.  For
each case the file stores the inputs (sample embedding matrix, worker count, mini batch, snapshot
capacity, epochs, batches per epoch) and the reference's outputs for every batch and worker: the
communication plan (as ascending keys; the reference's is an unordered set) and the sample
distribution.  Ragged lists are stored concatenated with offsets."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import laia_ref  # noqa: E402

#        W  mini  T  batches/epoch  capacity epochs  vocab  zipf
CASES = [(2, 4, 3, 3, 10, 1, 50, 1.3),
         (4, 8, 5, 4, 40, 2, 200, 1.2),
         (8, 16, 26, 3, 300, 1, 5000, 1.05),
         (3, 5, 2, 5, 6, 2, 30, 1.5),
         (1, 8, 4, 3, 20, 1, 40, 1.3),
         (4, 32, 26, 2, 2000, 1, 100000, 1.05)]


def main():
    out = {"ncases": np.int64(len(CASES))}
    for c, (W, mini, T, nb, cap, ep, vocab, a) in enumerate(CASES):
        rng = np.random.default_rng(100 + c)
        S = W * mini * nb
        emb = ((rng.zipf(a, (S, T)) - 1) % vocab).astype(np.int32)
        res = laia_ref.run(emb, ep, mini, nb, W, cap)
        plan_flat, plan_off, dist = [], [0], []
        for plans, d in res:
            for w in range(W):
                plan_flat.extend(plans[w])
                plan_off.append(len(plan_flat))
            dist.append(d)
        out["c%d_params" % c] = np.array([W, mini, T, nb, cap, ep], np.int64)
        out["c%d_emb" % c] = emb
        out["c%d_plan" % c] = np.array(plan_flat, np.int64)
        out["c%d_plan_off" % c] = np.array(plan_off, np.int64)
        out["c%d_dist" % c] = np.array(dist, np.int64)          # [batches, W, mini]
        print("case", c, "batches", len(res), "plan keys", len(plan_flat))
    np.savez_compressed(os.path.join(HERE, "laia_cases.npz"), **out)


if __name__ == "__main__":
    main()
