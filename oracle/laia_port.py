"""TEST INFRASTRUCTURE: pure-Python restatement of Herald's "Laia" embedding scheduler.

Restates, for small cases, the planner of the reference:

* ``MiniLRUCache``   laia/include/mini_lru_cache.h:14-137 (same semantics as the older
  python/hetu/laia/MiniLRUCache.h the Cython planner uses): an LRU of keys with a valid bit;
  ``get`` return codes -1 hit, -2 stale hit, 0 miss, 1 miss that evicted a valid line
  (mini_lru_cache.h:69-105).
* ``LaiaPlanner``    laia/src/laia_scheduler.cc:115-169 (the launch loop: epochs, the extra batch
  of the last epoch, the snapshot update) and :171-271 (get_dist: scoring, greedy assignment
  with the rotating tie-break ``(j + batch_id) % W``, communication plan).

Pinned against the reference's own Cython planner (python/hetu/laia/laia.pyx, built as
oracle/_ref/laia*.so by oracle/Makefile) by tests/test_laia.py and by the golden fixtures
tests/golden/laia_*.npz generated from it (tests/golden/make_golden_laia.py).  The C++ planner of
the reference (laia/src/*.cc) needs Boost and cannot be built here; laia.pyx is the same
algorithm (its plan is an unordered set where the C++ one is a sorted flat_set: plans are
compared as sorted key lists).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from collections import OrderedDict


class MiniLRUCache(object):
    """mini_lru_cache.h:14-137 without the bitmap (the scheduler default-constructs it)."""

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.lines = OrderedDict()      # key -> valid; last = most recent (list_.front())

    def check(self, key):               # :55-63
        return self.lines.get(key, False)

    def get(self, key):                 # :69-85
        if key not in self.lines:
            return self.insert(key)
        res = -1 if self.lines[key] else -2
        self.lines.move_to_end(key)
        self.lines[key] = True
        return res

    def insert(self, key):              # :88-105
        self.lines[key] = True
        self.lines.move_to_end(key)
        if len(self.lines) > self.capacity:
            _, valid = self.lines.popitem(last=False)
            return 1 if valid else 0
        return 0

    def outdate(self, key):             # :118-125
        if key in self.lines:
            self.lines[key] = False

    def evict(self, key):               # :107-116
        self.lines.pop(key, None)

    def get_keys(self):                 # :127-136: valid keys, ascending
        return sorted(k for k, v in self.lines.items() if v)


class LaiaPlanner(object):
    """One instance = the scheduler every worker runs (identical on every rank; `rank` only selects
    what is handed out).  ``next_all()`` returns ``(plans, dist)`` for all workers of the next batch,
    or None after the last one."""

    def __init__(self, sample_embs, mini_batch_size, nrank, cache_size, epoch_num=1, batch_num=1):
        self.embs = [[int(x) for x in row] for row in sample_embs]
        self.num_sample = len(self.embs)
        self.num_table = len(self.embs[0]) if self.embs else 0
        self.W = int(nrank)
        self.mini = int(mini_batch_size)
        self.batch_size = self.mini * self.W                      # laia_scheduler.cc:47
        self.snaps = [MiniLRUCache(cache_size) for _ in range(self.W)]
        self.epoch_num, self.batch_num = int(epoch_num), int(batch_num)
        self.epoch_id, self.batch_id = 0, 0
        self.in_epoch = False

    def _advance(self):
        """laia_scheduler.cc:126-135, 166: epochs x batches, one more batch in the last epoch."""
        while True:
            if not self.in_epoch:
                if self.epoch_id >= self.epoch_num:
                    return False
                self.epoch_id += 1
                self.batch_id = 0
                if self.epoch_id == self.epoch_num:
                    self.batch_num += 1
                self.in_epoch = True
            if self.batch_id < self.batch_num:
                return True
            self.in_epoch = False

    def get_dist(self):
        """laia_scheduler.cc:171-271."""
        W, S = self.W, self.num_sample
        start = (self.batch_id * self.batch_size) % S
        pos = [(start + i) % S for i in range(self.batch_size)]
        scores = [[0] * W for _ in pos]
        dep = [[[] for _ in range(W)] for _ in pos]
        for i, p in enumerate(pos):                               # :194-231 scoring
            for emb in self.embs[p]:
                for z in range(W):
                    if self.snaps[z].check(emb):
                        scores[i][z] += 1
                        dep[i][z].append(emb)
        workload = [0] * W
        dist = [[0] * self.mini for _ in range(W)]
        dist_keys = [set() for _ in range(W)]
        for i, p in enumerate(pos):                               # :233-254 greedy assignment
            best, best_w = -1, -1
            for j in range(W):
                w = (j + self.batch_id) % W
                if workload[w] < self.mini and best < scores[i][w]:
                    best, best_w = scores[i][w], w
            dist[best_w][workload[best_w]] = p
            dist_keys[best_w].add(p)
            workload[best_w] += 1
        plans = []
        for w in range(W):                                        # :256-270 communication plan
            plan = set()
            for i, p in enumerate(pos):
                if p not in dist_keys[w]:
                    plan.update(dep[i][w])
            plans.append(sorted(plan))
        return plans, dist

    def next_all(self):
        if not self._advance():
            return None
        plans, dist = self.get_dist()
        for w in range(self.W):                                   # :146-161 snapshot update
            for key in plans[w]:
                self.snaps[w].outdate(key)
            uniq = sorted({e for p in dist[w] for e in self.embs[p]})
            for key in uniq:
                self.snaps[w].get(key)
        self.batch_id += 1
        return plans, dist
