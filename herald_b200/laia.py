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


class _ShmRing(object):
    """One worker's message ring in POSIX shared memory (laia/include/share_mem.h:39-160)."""

    def __init__(self, name, create, data_bytes=1 << 30, wait_s=120.0):
        import time
        self._h = ctypes.c_void_p()
        deadline = time.time() + (0 if create else wait_s)
        while True:
            # a reader waits for the creating rank instead of betting on the reference's fixed
            # 3 s head start (laia_dataloader.py:77-78)
            if _LIB.hb_shmring_open(ctypes.byref(self._h), name.encode(), int(create), _sz(data_bytes)) == 0:
                break
            if time.time() >= deadline:
                check_call(-1)
            time.sleep(0.05)

    def send(self, words):
        """Blocks (polling, as the reference does: topk_scheduler.cc:229-242) until there is room."""
        import time
        arr = np.ascontiguousarray(words, dtype=np.uint64)
        sent = ctypes.c_longlong()
        while True:
            check_call(_LIB.hb_shmring_send(self._h, arr.ctypes.data_as(ctypes.c_void_p), _sz(arr.size),
                                            ctypes.byref(sent)))
            if sent.value >= 0:
                return
            time.sleep(1e-5)

    def recv(self, timeout_s=None):
        """Blocks until a message is there (topk_scheduler.cc:254-278); the reference waits for
        ever, here a planner that died is reported after `timeout_s` ($HERALD_LAIA_TIMEOUT_S, 600)."""
        import os
        import time
        if timeout_s is None:
            timeout_s = float(os.environ.get("HERALD_LAIA_TIMEOUT_S", "600"))
        deadline = time.time() + timeout_s
        n = ctypes.c_longlong()
        while True:
            if time.time() > deadline:
                raise RuntimeError("no message from the planning process within %.0f s" % timeout_s)
            check_call(_LIB.hb_shmring_recv(self._h, None, _sz(0), ctypes.byref(n)))
            if n.value >= 0:
                out = np.empty(max(n.value, 1), np.uint64)
                check_call(_LIB.hb_shmring_recv(self._h, out.ctypes.data_as(ctypes.c_void_p), _sz(out.size),
                                                ctypes.byref(n)))
                return out[:n.value].tolist()
            time.sleep(1e-5)

    def used(self):
        w = _sz(0)
        check_call(_LIB.hb_shmring_used(self._h, ctypes.byref(w)))
        return w.value

    def close(self):
        if self._h:
            _LIB.hb_shmring_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TopkScheduler(LaiaScheduler):
    """The reference's production planner (laia/src/topk_scheduler.cc, bound at
    laia/src/python_binding.cc:16-22): scores over the dataset's pre-profiled top-k tables, per-thread
    slot split, and the `local_shared` mode in which local rank 0 of a node plans for all
    `local_size` local workers and hands each its messages through a shared-memory ring
    ("laia_cache_<local rank>"); the others only read theirs (`pop_from_local_worker`)."""

    def __init__(self):
        super().__init__()
        self.local_shared = False
        self._rings, self._my_ring, self._thread = [], None, None

    def start(self, sample_embs, num_sample, num_table, epoch_num, mini_batch_size, batch_num, nrank, rank,
              cache_size, num_threads, dataset, top_k_table, local_shared=False, local_rank=0, local_size=1,
              ring_bytes=1 << 30):
        self.local_shared, self.local_rank, self.local_size = bool(local_shared), int(local_rank), int(local_size)
        self.rank, self.nrank, self.mini_batch_size = int(rank), int(nrank), int(mini_batch_size)
        self._queue, self._done = [], False
        if self.local_shared:
            # "laia_cache_<i>" in the reference (topk_scheduler.cc:70-80); here the name also carries
            # the job (the launcher's pid, common to the node's workers, or $HERALD_LAIA_SESSION), so a
            # ring left in /dev/shm by a crashed run is never mistaken for this run's
            import os
            job = os.environ.get("HERALD_LAIA_SESSION", str(os.getppid()))
            name = lambda i: "laia_cache_%s_%d" % (job, i)
            if self.local_rank == 0:        # creates every local ring, then opens its own
                self._rings = [_ShmRing(name(i), True, ring_bytes) for i in range(self.local_size)]
            self._my_ring = _ShmRing(name(self.local_rank), False)
            if self.local_rank != 0:
                return                      # no planner in this process (:181-185)
        embs = np.ascontiguousarray(sample_embs, dtype=np.uint64)
        if embs.ndim != 2:
            raise RuntimeError("Input should be 2D numpy array")
        assert embs.shape == (num_sample, num_table)
        self.close_planner()
        h = ctypes.c_void_p()
        check_call(_LIB.hb_laia_create_topk(ctypes.byref(h), embs.ctypes.data_as(ctypes.c_void_p),
                                            _sz(num_sample), _sz(num_table), _sz(epoch_num),
                                            _sz(mini_batch_size), _sz(batch_num), _sz(nrank), _sz(rank),
                                            _sz(cache_size), _sz(num_threads), str(dataset).encode(),
                                            _sz(top_k_table)))
        self._h = h
        if self.local_shared:
            import threading
            self._thread = threading.Thread(target=self._serve_local, daemon=True)
            self._thread.start()

    def _serve_local(self):
        """Plan batch after batch and push every local worker its plan and its sample positions
        (topk_scheduler.cc:291-296: worker id = this rank + local index), then the terminator."""
        while self.step():
            for i, ring in enumerate(self._rings):
                ring.send(self.plan_of(self.rank + i))
                ring.send(self.dist_of(self.rank + i))
        for ring in self._rings:
            ring.send([0])

    def pop(self):
        if self.local_shared:
            raise RuntimeError("local_shared: read with pop_from_local_worker()")
        return super().pop()

    def pop_from_local_worker(self):
        assert self.local_shared
        return self._my_ring.recv()

    def length(self):
        if self.local_shared:
            return self._my_ring.used()     # words waiting, as SharedMemBuf::queue_length (share_mem.h:163-165)
        return super().length()

    def close_planner(self):
        LaiaScheduler.close(self)

    def close(self):
        if self._thread is not None:
            self._thread.join(timeout=30)
            self._thread = None
        self.close_planner()
        if self._my_ring is not None:
            self._my_ring.close()
            self._my_ring = None
        for r in self._rings:
            r.close()
        self._rings = []


# top-k table counts the reference passes per dataset (python/hetu/laia/laia_dataloader.py:19-24)
top_k_table = {"criteo": 20, "avazu": 17, "movie": 2, "criteosearch": 16}
local_worker_num = 8


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
    `config` supplies `nrank`, `rank`, `local_rank`, `cache_limit` (HetuConfig).  `local_shared`
    selects the TopkScheduler over shared-memory rings (run_laia.py --local-shared)."""

    queue_size = 5

    def __init__(self, sparse_data, batch_size, drop_last=True, dataset="criteo", local_shared=False):
        # ids arrive float32-carried (python/hetu/dataloader.py:14); the planner wants integers
        self.sparse_data = np.asarray(sparse_data, np.float32).astype(np.intc)
        self.batch_size, self.drop_last, self.dataset = batch_size, drop_last, dataset
        self.local_shared = bool(local_shared)
        self.init = False

    def start(self, config, dataset_num=3, epoch_num=-1):
        assert not self.init, "LAIA scheduler can only be initialized once"
        self.local_rank = config.local_rank
        self.samples_num = len(self.sparse_data) // config.nrank
        self.batch_size = min(int(self.batch_size), self.samples_num // self.queue_size)
        assert self.batch_size > 0, "Batch size %d invalid." % self.batch_size
        whole, rest = divmod(self.samples_num, self.batch_size)
        self.batch_num = whole if (self.drop_last or rest == 0) else whole + 1
        epochs = epoch_num if epoch_num >= 0 else (1 << 62)               # -1: until closed
        if not self.local_shared:
            self.sched = LaiaScheduler()
            self.sched.start(self.sparse_data, self.sparse_data.shape[0], self.sparse_data.shape[1],
                             epochs, self.batch_size, self.batch_num, int(config.nrank), int(config.rank),
                             int(config.cache_limit), 16, 24)
        else:
            # laia_dataloader.py:72-96: local rank 0 plans for the node's workers (TopkScheduler, 80
            # threads, the dataset's pre-profiled top-k tables); the others wait for its rings
            self.sched = TopkScheduler()     # (a reader waits for its ring to appear: _ShmRing)
            self.sched.start(self.sparse_data, self.sparse_data.shape[0], self.sparse_data.shape[1],
                             epochs, self.batch_size, self.batch_num, int(config.nrank), int(config.rank),
                             int(config.cache_limit), int(getattr(config, "laia_threads", 80)), self.dataset,
                             int(top_k_table[self.dataset]), True, int(config.local_rank),
                             int(getattr(config, "local_size", local_worker_num)))
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
        msg = self.sched.pop_from_local_worker() if self.local_shared else self.sched.pop()
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


class LAIADataloader(object):
    """The "train" loader of a Laia run (python/hetu/laia/laia_dataloader.py:172-231): WHICH samples
    make up batch b is the planner's decision (`sched.get_input_index(b)`), and a sparse loader hands
    out the communication plan next to the ids — `(ids, plan)`, which CacheSparseTable /
    ParameterServerCommunicateOp route to `embedding_update_with_push_keys`.  Every loader of a
    step (sparse ids, labels, dense features) shares one LAIAScheduler and reports its progress
    with its `sched_id`."""

    def __init__(self, sched, sched_id, is_sparse, raw_data, batch_size, name="default", func=None,
                 drop_last=True):
        from . import ndarray
        self._nd = ndarray
        self.func = func if func else (lambda x: x)
        self.raw_data = np.array(self.func(raw_data), np.float32)
        self.sched, self.sched_id, self.is_sparse = sched, sched_id, is_sparse
        self.batch_size, self.name, self.drop_last = batch_size, str(name), drop_last

    def init_states(self, rank=None, nrank=None):
        self.samples_num, self.batch_num = self.sched.samples_num, self.sched.batch_num
        self.batch_size = self.sched.batch_size
        self.batch_index, self.rank = 0, rank

    def _get_arr(self, batch):
        idx = np.asarray(self.sched.get_input_index(batch), np.int64)
        rows = self._nd.array(self.raw_data[idx], ctx=self._nd.cpu(0))
        if not self.is_sparse:
            return rows
        # (the reference wraps the plan in a float32 NDArray, :198-203; it stays integer here: ids
        # above 2^24 would not survive the cast, and the cache wants uint64 push keys anyway)
        return rows, np.asarray(self.sched.get_comm_plan(batch), np.uint64)

    def get_arr(self):
        res = self._get_arr(self.batch_index)
        self.batch_index = (self.batch_index + 1) % self.batch_num
        self.sched.step_forward(self.sched_id)      # after the reads of this batch (:150)
        return res

    def get_next_arr(self):
        return self._get_arr(self.batch_index)

    def get_cur_shape(self):
        return (len(self.sched.get_input_index(self.batch_index)),) + tuple(self.raw_data.shape[1:])


def laia_dataloader_op(dataloaders, sched, sched_id, is_sparse=False):
    """laia_dataloader.py:234-259: the loader named "train" follows the planner, the others are
    plain Dataloaders."""
    from .dataloader import Dataloader, DataloaderOp
    built = []
    for dl in dataloaders:
        if isinstance(dl, (Dataloader, LAIADataloader)):
            built.append(dl)
        elif isinstance(dl, list):
            if len(dl) >= 3 and dl[2] == "train":
                built.append(LAIADataloader(sched, sched_id, is_sparse, *dl))
            else:
                built.append(Dataloader(*dl))
        elif isinstance(dl, dict):
            if dl.get("name") == "train":
                built.append(LAIADataloader(sched=sched, sched_id=sched_id, is_sparse=is_sparse, **dl))
            else:
                built.append(Dataloader(**dl))
        else:
            raise AssertionError("Dataloader parameter invalid.")
    return DataloaderOp(built)


# top-k table orders the reference pre-profiled per dataset (laia/src/topk_scheduler.cc:150-165)
TOPK_TABLE_ORDER = {
    "criteo": [9, 13, 22, 20, 12, 21, 17, 14, 24, 3, 5, 10, 16, 15, 19, 2, 4, 11, 7, 25, 23, 18, 8, 1, 0, 6],
    "avazu": [1, 2, 4, 5, 15, 7, 6, 16, 12, 0, 17, 8, 14, 10, 9, 11, 13, 3],
    "movie": [0, 1],
    "criteosearch": [0, 11, 3, 4, 5, 14, 1, 6, 2, 13, 16, 9, 8, 10, 12, 7, 15],
}


def assign_by_scores(scores, mini_batch_size, parts=1):
    """The TopkScheduler assignment (laia/src/topk_scheduler.cc:430-452) on a finished score matrix
    `scores[W, S]` (S = W * mini_batch_size samples of one global batch): logical thread t owns a
    contiguous run of samples and, of every worker's mini-batch, a contiguous run of slots; a sample
    goes to the best-scoring worker that still has room in the thread's slots, scanning from the
    candidate (here: the lowest-ranked worker holding the maximum) and stopping at it if it has room.
    -> int64[W, mini_batch_size] sample positions inside the batch.  Deterministic: every rank that
    holds the same matrix computes the same assignment."""
    scores = np.asarray(scores)
    W, S = scores.shape
    mini = int(mini_batch_size)
    assert S == W * mini and mini % parts == 0
    dist = np.zeros((W, mini), np.int64)
    per, room = S // parts, mini // parts
    for t in range(parts):
        load = [0] * W
        for i in range(t * per, (t + 1) * per):
            col = scores[:, i]
            cand = int(np.argmax(col))
            best, best_w = -1, -1
            for j in range(W):
                w = (j + cand) % W
                if best < int(col[w]) and load[w] < room:
                    best, best_w = int(col[w]), w
                    if w == cand:
                        break
            dist[best_w, t * room + load[best_w]] = i
            load[best_w] += 1
    return dist


class GpuScoredPlanner(object):
    """Herald planning against the REAL caches instead of simulated snapshots: every
    rank scores the whole global batch against its own cache index on the GPU
    (`hb_cache_score`), the W score columns are all-gathered, every rank runs the same assignment,
    and a rank's communication plan is the set of ids of ITS samples that it already holds
    (`hb_cache_probe`) — the TopkScheduler's rule (topk_scheduler.cc:476-483) without its
    erase-while-iterating artefact.  `allgather(column) -> [W, S]` is supplied by the launcher
    (torch.distributed / NCCL); None for a single worker."""

    def __init__(self, cache, sample_embs, mini_batch_size, nrank, rank, dataset="criteo", top_k=None,
                 parts=1, allgather=None, fresh=True):
        self.cache = cache
        self.embs = np.ascontiguousarray(sample_embs, dtype=np.uint64)
        self.mini, self.W, self.rank, self.parts = int(mini_batch_size), int(nrank), int(rank), int(parts)
        order = [t for t in TOPK_TABLE_ORDER[dataset] if t < self.embs.shape[1]]
        k = len(order) if top_k is None else min(int(top_k), len(order))
        self.order = np.asarray(order[:k], np.uint32)
        self.allgather, self.fresh = allgather, fresh

    def plan_batch(self, b):
        """-> (sample positions of this rank in the sample matrix, ascending plan keys)."""
        B, S = self.mini * self.W, self.embs.shape[0]
        pos = (b * B + np.arange(B)) % S
        ids = self.embs[pos]
        mine = self.cache.score_samples(ids, self.order, len(self.order), fresh=self.fresh)
        cols = mine[None, :] if self.allgather is None else np.asarray(self.allgather(mine))
        assert cols.shape == (self.W, B)
        dist = assign_by_scores(cols, self.mini, self.parts)
        my_pos = pos[dist[self.rank]]
        keys = np.unique(self.embs[my_pos].reshape(-1))
        return my_pos, keys[self.cache.resident(keys)]
