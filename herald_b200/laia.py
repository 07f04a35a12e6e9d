"""Laia / Herald embedding scheduler — the counterpart of the reference's pybind module
``laia_cache`` (laia/src/python_binding.cc:8-16) and of python/hetu/laia/laia_dataloader.py's use
of it.

``LaiaScheduler().start(...)`` then ``pop()`` hands out, batch after batch, first the
communication plan of this rank (list of embedding ids it must push before the others use them),
then the sample indices this rank trains on — the same two-message wire format as the reference
(laia/src/laia_scheduler.cc:138-139) — and finally ``[0]`` (:168).  The planner itself is C++ in
libherald_b200.so (csrc/hb_laia.cu); a batch is planned when its first message is popped.
"""
import ctypes

import numpy as np

from ._base import _LIB, check_call

_sz = ctypes.c_size_t


class LaiaScheduler(object):
    def __init__(self):
        self._h = ctypes.c_void_p()
        self._queue = []
        self._done = False
        self.rank = 0
        self.nrank = 1
        self.mini_batch_size = 0

    # LaiaScheduler::start(sample_embs_py, num_sample, num_table, epoch_num, mini_batch_size,
    #                      batch_num, nrank, rank, cache_size, num_threads, top_k_table)
    def start(self, sample_embs, num_sample, num_table, epoch_num, mini_batch_size, batch_num,
              nrank, rank, cache_size, num_threads=8, top_k_table=0):
        embs = np.ascontiguousarray(sample_embs, dtype=np.uint64)
        if embs.ndim != 2:
            raise RuntimeError("Input should be 2D numpy array")     # laia_scheduler.cc:35-36
        assert embs.shape == (num_sample, num_table)
        self.close()
        h = ctypes.c_void_p()
        check_call(_LIB.hb_laia_create(ctypes.byref(h), embs.ctypes.data_as(ctypes.c_void_p),
                                       _sz(num_sample), _sz(num_table), _sz(epoch_num),
                                       _sz(mini_batch_size), _sz(batch_num), _sz(nrank), _sz(rank),
                                       _sz(cache_size), _sz(num_threads)))
        self._h = h
        self._queue, self._done = [], False
        self.rank, self.nrank, self.mini_batch_size = int(rank), int(nrank), int(mini_batch_size)

    def _plan_next(self):
        done = ctypes.c_int(0)
        check_call(_LIB.hb_laia_next(self._h, ctypes.byref(done)))
        if done.value:
            self._done = True
            self._queue.append([0])                                   # laia_scheduler.cc:168
            return
        plan, dist = self.plan_of(self.rank), self.dist_of(self.rank)
        self._queue.append(plan.tolist())
        self._queue.append(dist.tolist())

    def pop(self):
        if not self._queue:
            if self._done:
                raise RuntimeError("the scheduler has finished")
            self._plan_next()
        return self._queue.pop(0)

    def length(self):
        """Messages ready to pop; batches are planned on demand, so an unfinished scheduler
        always has the next pair available (laia_dataloader.py:161-163 only asks for >= 2)."""
        return len(self._queue) if self._done else max(2, len(self._queue))

    # --- every worker's part of the most recent batch (tests, a multi-rank driver) ---
    def plan_of(self, worker):
        n = _sz(0)
        check_call(_LIB.hb_laia_plan_size(self._h, _sz(worker), ctypes.byref(n)))
        out = np.empty(n.value, np.uint64)
        check_call(_LIB.hb_laia_plan(self._h, _sz(worker), out.ctypes.data_as(ctypes.c_void_p), n))
        return out

    def dist_of(self, worker):
        out = np.empty(self.mini_batch_size, np.uint64)
        check_call(_LIB.hb_laia_dist(self._h, _sz(worker), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def snapshot_keys(self, worker):
        n = _sz(0)
        check_call(_LIB.hb_laia_snapshot_keys(self._h, _sz(worker), None, _sz(0), ctypes.byref(n)))
        out = np.empty(n.value, np.uint64)
        check_call(_LIB.hb_laia_snapshot_keys(self._h, _sz(worker), out.ctypes.data_as(ctypes.c_void_p),
                                              n, ctypes.byref(n)))
        return out

    def step(self):
        """Plan one batch without the queue; False when the sequence is over."""
        done = ctypes.c_int(0)
        check_call(_LIB.hb_laia_next(self._h, ctypes.byref(done)))
        return not done.value

    def close(self):
        if self._h:
            _LIB.hb_laia_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MiniLRUCache(object):
    """laia/include/mini_lru_cache.h:14-137 (the planner's simulated worker cache)."""

    def __init__(self, capacity):
        self._h = ctypes.c_void_p()
        check_call(_LIB.hb_minilru_create(ctypes.byref(self._h), _sz(capacity)))

    def get(self, key):
        return int(_LIB.hb_minilru_get(self._h, ctypes.c_uint64(int(key))))

    def check(self, key):
        return bool(_LIB.hb_minilru_check(self._h, ctypes.c_uint64(int(key))))

    def outdate(self, key):
        _LIB.hb_minilru_outdate(self._h, ctypes.c_uint64(int(key)))

    def evict(self, key):
        _LIB.hb_minilru_evict(self._h, ctypes.c_uint64(int(key)))

    def get_keys(self):
        n = _sz(0)
        check_call(_LIB.hb_minilru_keys(self._h, None, _sz(0), ctypes.byref(n)))
        out = np.empty(n.value, np.uint64)
        check_call(_LIB.hb_minilru_keys(self._h, out.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(n)))
        return out

    def __del__(self):
        try:
            if self._h:
                _LIB.hb_minilru_destroy(self._h)
        except Exception:
            pass
