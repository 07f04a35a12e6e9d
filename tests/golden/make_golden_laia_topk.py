"""Generate tests/golden/laia_topk_cases.npz from the REFERENCE's TopkScheduler.

Run in the authoring container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_laia_topk.py
oracle/_ref/laia_cache*.so is laia/src/topk_scheduler.cc (+ thread_pool / array / utils)
compiled unmodified.  Per case: the inputs (sample matrix, worker count, mini batch, snapshot
capacity, epochs, batches per epoch, planner threads, dataset name, top-k) and the reference's
outputs for every batch and worker (communication plan, sample distribution), ragged lists stored
concatenated with offsets."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import laia_cpp_ref  # noqa: E402

#        W  mini  T  batches  capacity epochs threads dataset  topk  vocab  zipf
CASES = [(2, 4, 26, 3, 40, 1, 2, "criteo", 20, 300, 1.3),
         (4, 16, 26, 4, 300, 2, 4, "criteo", 20, 3000, 1.1),
         (8, 16, 26, 3, 500, 1, 8, "criteo", 26, 5000, 1.05),
         (3, 6, 18, 5, 30, 1, 3, "avazu", 17, 100, 1.4),
         (4, 8, 2, 4, 10, 2, 1, "movie", 2, 40, 1.3),
         (1, 8, 17, 3, 50, 1, 4, "criteosearch", 16, 200, 1.2),
         (4, 32, 26, 2, 4000, 1, 16, "criteo", 0, 100000, 1.05)]


def main():
    out = {"ncases": np.int64(len(CASES))}
    for c, (W, mini, T, nb, cap, ep, th, ds, topk, vocab, a) in enumerate(CASES):
        rng = np.random.default_rng(300 + c)
        S = W * mini * nb
        emb = ((rng.zipf(a, (S, T)) - 1) % vocab + 1).astype(np.int64)     # ids >= 1: [0] ends the wire
        res = laia_cpp_ref.run_topk(emb, ep, mini, nb, W, cap, th, ds, topk)
        plan_flat, plan_off, dist = [], [0], []
        for plans, d in res:
            for w in range(W):
                plan_flat.extend(plans[w])
                plan_off.append(len(plan_flat))
            dist.append(d)
        out["c%d_params" % c] = np.array([W, mini, T, nb, cap, ep, th, topk], np.int64)
        out["c%d_dataset" % c] = np.array(ds)
        out["c%d_emb" % c] = emb
        out["c%d_plan" % c] = np.array(plan_flat, np.int64)
        out["c%d_plan_off" % c] = np.array(plan_off, np.int64)
        out["c%d_dist" % c] = np.array(dist, np.int64)          # [batches, W, mini]
        print("case", c, "batches", len(res), "plan keys", len(plan_flat))
    np.savez_compressed(os.path.join(HERE, "laia_topk_cases.npz"), **out)


if __name__ == "__main__":
    main()
