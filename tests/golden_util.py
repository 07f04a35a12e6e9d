"""Replay a golden fixture (tests/golden/*.npz, generated from the reference by
tests/golden/make_golden.py) against any cache implementation with the harness surface."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PULL = ("num_all", "num_unique", "num_miss", "num_transfered", "is_full")
PUSH = ("num_all", "num_unique", "num_miss", "num_evict", "num_transfered", "is_full")


def fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "cache_*.npz")))


def parse_name(path):
    _, policy, limit, bound, mode = os.path.basename(path)[:-4].split("_")
    return policy, int(limit), int(bound), mode


def replay(path, make_cache, bits_equal):
    """make_cache(policy, limit, bound, rows0) -> object with
         lookup(keys)->dest, update(keys, grads, push_keys)->perf dict, push_pull(k1,k2,g)->dest,
         last_pull_perf(), keys(), rows(), versions(), line(key)->(version, data)"""
    z = np.load(path)
    policy, limit, bound, mode = parse_name(path)
    c = make_cache(policy, limit, bound, z["rows0"])

    def seg(name, i):
        off = z[name + "_off"]
        return z[name][off[i]:off[i + 1]]

    for i, kind in enumerate(z["kinds"]):
        tag = "%s call %d kind %d" % (os.path.basename(path), i, kind)
        keys = seg("keys", i)
        if kind == 0:
            dest = c.lookup(keys)
            bits_equal(dest, seg("dest", i), tag + " dest")
            got = c.last_perf()
            exp = dict(zip(PULL, z["perf"][i][:5]))
            for k in PULL:
                assert int(got[k]) == int(exp[k]), (tag, k, got, exp)
        elif kind in (1, 2):
            pk = seg("push_keys", i) if kind == 2 else None
            c.update(keys, seg("grads", i), pk)
            got = c.last_perf()
            exp = dict(zip(PUSH, z["perf"][i]))
            for k in PUSH:
                assert int(got[k]) == int(exp[k]), (tag, k, got, exp)
        else:
            dest = c.push_pull(keys, seg("push_keys", i), seg("grads", i))
            bits_equal(dest, seg("dest", i), tag + " push_pull dest")
    assert np.array_equal(c.keys(), z["resident_keys"]), path
    bits_equal(c.rows(), z["final_rows"], path + " final rows")
    assert np.array_equal(c.versions(), z["final_versions"]), path
    for k, ver, data in zip(z["resident_keys"], z["line_versions"], z["line_data"]):
        v, d = c.line(int(k))
        assert v == ver, (path, int(k), v, ver)
        bits_equal(d, data, "%s line %d" % (path, int(k)))
