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


class LAIAScheduler(object):
    """Per-process front end of the planner, with the surface `run_laia.py` uses from
    python/hetu/laia/laia_dataloader.py:28-169: `LAIAScheduler(sparse_data, batch_size)`,
    `start(config)`, `get_input_index(b)`, `get_comm_plan(b)`, `step_forward(dataset_id)`.

    It keeps `queue_size` (5) batches planned ahead.  What it serves for batch b is the pair
    (sample indices of b, communication plan computed for b + 1): the very first plan is dropped
    (:108-114), so the keys a worker pushes after training b are the ones its peers need in b + 1.
    A slot is refilled once every dataset (train / validate / ...) has stepped past it (:150-169).
    `config` supplies `nrank`, `rank`, `local_rank`, `cache_limit` (HetuConfig).  The reference's
    `local_shared` mode (TopkScheduler over shared memory) is not provided."""

    queue_size = 5

    def __init__(self, sparse_data, batch_size, drop_last=True, dataset="criteo", local_shared=False):
        if local_shared:
            raise NotImplementedError("the TopkScheduler (local_shared) variant is not provided")
        # ids arrive float32-carried (python/hetu/dataloader.py:14); the planner wants integers
        self.sparse_data = np.asarray(sparse_data, np.float32).astype(np.intc)
        self.batch_size, self.drop_last, self.dataset = batch_size, drop_last, dataset
        self.local_shared = False
        self.init = False

    def start(self, config, dataset_num=3, epoch_num=-1):
        assert not self.init, "LAIA scheduler can only be initialized once"
        self.local_rank = config.local_rank
        self.samples_num = len(self.sparse_data) // config.nrank
        self.batch_size = min(int(self.batch_size), self.samples_num // self.queue_size)
        assert self.batch_size > 0, "Batch size %d invalid." % self.batch_size
        whole, rest = divmod(self.samples_num, self.batch_size)
        self.batch_num = whole if (self.drop_last or rest == 0) else whole + 1
        self.sched = LaiaScheduler()
        self.sched.start(self.sparse_data, self.sparse_data.shape[0], self.sparse_data.shape[1],
                         epoch_num if epoch_num >= 0 else (1 << 62),       # -1: until closed
                         self.batch_size, self.batch_num, int(config.nrank), int(config.rank),
                         int(config.cache_limit), 16, 24)
        self.channel_close = False
        self._receive()                                   # the plan of batch 0: nothing was cached yet
        self._slots = {b: self._receive_pair() for b in range(self.queue_size)}
        self.step = [0] * dataset_num
        self.cur_min_step = 0
        self.init = True

    def _receive(self):
        """One message of the planner's wire (plan keys or sample indices); [] once it has ended.
        As in the reference (:137-139) a message that is exactly [0] ends the channel."""
        if self.channel_close:
            raise RuntimeError("the scheduler channel is closed")
        msg = self.sched.pop()
        assert isinstance(msg, list)
        if msg == [0]:
            self.channel_close = True
            return []
        return msg

    def _receive_pair(self):
        indices = self._receive()
        return indices, self._receive()

    def get_input_index(self, batch_id):
        return self._slots[batch_id][0]

    def get_comm_plan(self, batch_id):
        return self._slots[batch_id][1]

    def step_forward(self, dataset_id):
        """NOTE (as in the reference): call after get_input_index / get_comm_plan of the batch."""
        self.step[dataset_id] += 1
        slowest = min(self.step)
        while self.cur_min_step < slowest and not self.channel_close:
            if self.sched.length() < 2 and slowest - self.cur_min_step < self.queue_size:
                break                                     # nothing planned yet and no need to wait
            done = self.cur_min_step % self.batch_num
            del self._slots[done]
            self._slots[(done + self.queue_size) % self.batch_num] = self._receive_pair()
            self.cur_min_step += 1
