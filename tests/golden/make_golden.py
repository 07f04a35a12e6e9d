"""Generate the golden vectors under tests/golden/ from the REFERENCE ITSELF.

Run in the authoring container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
It drives oracle/_ref — the reference's src/hetu_cache + ps-lite/src/PSFhandle_embedding.cc
compiled unmodified, transport replaced by oracle/ref_shim/transport_shim.cc — with seeded call
sequences and stores inputs AND the reference's outputs.  The fixtures travel to the GPU box,
/root/reference does not.

Each cache_<policy>_<limit>_<bound>_<mode>.npz holds, for a sequence of steps:
  rows0                          initial table                                [V, D] f32
  step kinds / keys / grads / push_keys / pull2 / push2 / grads2 (ragged, concatenated + offsets)
  dest (concatenated gathered rows), perf counters per call
  final rows, versions, resident keys, per-line versions and data
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref  # noqa: E402

V, D, STEPS = 200, 8, 40
PULL = ("num_all", "num_unique", "num_miss", "num_transfered", "is_full")
PUSH = ("num_all", "num_unique", "num_miss", "num_evict", "num_transfered", "is_full")


def make(policy, limit, bound, mode, seed):
    rng = np.random.default_rng(seed)
    rows0 = rng.normal(0, 0.01, (V, D)).astype(np.float32)
    srv = ref.Server(V, D, rows0)
    c = ref.Cache(srv, policy, limit, bound)
    calls = []  # (kind, keys, grads, push_keys, dest, perf)
    for t in range(STEPS):
        n = int(rng.integers(1, 60))
        keys = ((rng.zipf(1.3, n) - 1) % V).astype(np.uint64)
        dest = c.embedding_lookup(keys)
        p = c.perf[-1]
        calls.append(dict(kind=0, keys=keys, dest=dest.copy(),
                          perf=[int(p[k]) for k in PULL] + [0]))
        ukeys = keys if rng.random() < 0.7 else ((rng.zipf(1.3, int(rng.integers(1, 60))) - 1) % V).astype(np.uint64)
        grads = rng.normal(0, 1e-3, (len(ukeys), D)).astype(np.float32)
        pk = None
        if mode == "plan":
            pk = np.unique(rng.choice(ukeys, size=max(1, len(ukeys) // 3))).astype(np.uint64)
        p = c.embedding_update(ukeys, grads, pk)
        calls.append(dict(kind=1 if pk is None else 2, keys=ukeys, grads=grads, push_keys=pk,
                          perf=[int(p[k]) for k in PUSH]))
        if mode == "pushpull" and t % 3 == 0:
            k1 = ((rng.zipf(1.3, n) - 1) % V).astype(np.uint64)
            k2 = ((rng.zipf(1.3, int(rng.integers(1, 60))) - 1) % V).astype(np.uint64)
            g2 = rng.normal(0, 1e-3, (len(k2), D)).astype(np.float32)
            dest = c.embedding_push_pull(k1, k2, g2)
            calls.append(dict(kind=3, keys=k1, dest=dest.copy(), push_keys=k2, grads=g2,
                              perf=[0] * 6))
    keys_res = c.keys()
    final_rows, final_ver = srv.rows(), srv.versions()
    line_ver = np.zeros(len(keys_res), np.int64)
    line_data = np.zeros((len(keys_res), D), np.float32)
    for i, k in enumerate(keys_res):
        ln = c.line(int(k))
        line_ver[i], line_data[i] = ln["version"], ln["data"]

    def cat(field, width=None, dtype=None):
        parts, offs = [], [0]
        for cl in calls:
            a = cl.get(field)
            if a is None:
                a = np.zeros((0,) if width is None else (0, width), dtype)
            parts.append(np.asarray(a, dtype))
            offs.append(offs[-1] + len(a))
        return np.concatenate(parts), np.asarray(offs, np.int64)

    keys_c, keys_o = cat("keys", None, np.uint64)
    grads_c, grads_o = cat("grads", D, np.float32)
    pk_c, pk_o = cat("push_keys", None, np.uint64)
    dest_c, dest_o = cat("dest", D, np.float32)
    out = os.path.join(HERE, "cache_%s_%d_%d_%s.npz" % (policy, limit, bound, mode))
    np.savez_compressed(
        out, rows0=rows0, kinds=np.array([cl["kind"] for cl in calls], np.int32),
        keys=keys_c, keys_off=keys_o, grads=grads_c, grads_off=grads_o, push_keys=pk_c,
        push_keys_off=pk_o, dest=dest_c, dest_off=dest_o,
        perf=np.array([cl["perf"] for cl in calls], np.int64), final_rows=final_rows,
        final_versions=final_ver, resident_keys=keys_res, line_versions=line_ver,
        line_data=line_data, meta=np.array([limit, bound, V, D], np.int64))
    return out


if __name__ == "__main__":
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    seed = 100
    for policy in ("lru", "lfu", "lfuopt"):
        for limit, bound, mode in ((5, 0, "plain"), (30, 2, "plain"), (150, 10, "plain"),
                                   (30, 10, "plan"), (30, 0, "pushpull")):
            seed += 1
            print(make(policy, limit, bound, mode, seed))
