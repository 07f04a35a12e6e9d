"""CacheSparseTable — mirrors python/hetu/cstable.py:19-248 (same constructor, method names,
dtype/shape asserts and perf helpers).  The one deliberate difference: NDArray arguments may sit
on the GPU as well as on the host (the reference asserts CPU because its cache is host code)."""
import numpy as np

from . import ndarray
from . import hetu_cache
from .ps import get_worker_communicate


class CacheSparseTable:
    def __init__(self, limit, length, width, node_id, policy="LRU", bound=100):
        comm = get_worker_communicate()
        policy = policy.lower()
        if policy == "lru":
            self.cache = hetu_cache.LRUCache(limit, length, width, node_id)
        elif policy == "lfu":
            self.cache = hetu_cache.LFUCache(limit, length, width, node_id)
        elif policy == "lfuopt":
            self.cache = hetu_cache.LFUOptCache(limit, length, width, node_id)
        else:
            raise NotImplementedError(policy)
        self.cache.pull_bound = bound
        self.cache.push_bound = bound
        comm.BarrierWorker()

    @staticmethod
    def _done(wait, sync):
        if sync:
            wait.wait()
        return wait

    def embedding_lookup(self, keys, dest, sync=False):
        if isinstance(keys, tuple):      # Laia dataloader hands (ids, comm_plan)
            keys = keys[0]
        if type(keys) is np.ndarray and type(dest) is np.ndarray:
            assert dest.shape == (keys.size, self.width)
            assert keys.dtype == np.uint64
            assert dest.dtype == np.float32
            wait = self.cache.embedding_lookup(keys, dest)
        elif type(keys) is ndarray.NDArray and type(dest) is ndarray.NDArray:
            assert dest.shape == (*keys.shape, self.width)
            wait = self.cache.embedding_lookup_raw(keys.data_ptr, dest.data_ptr,
                                                   int(np.prod(keys.shape)))
            wait._keep = (keys, dest)
        else:
            raise TypeError
        return self._done(wait, sync)

    def embedding_update(self, keys, grads, sync=False):
        if type(keys) is np.ndarray and type(grads) is np.ndarray:
            assert grads.shape == (keys.size, self.width)
            assert keys.dtype == np.uint64
            assert grads.dtype == np.float32
            wait = self.cache.embedding_update(keys, grads)
        elif type(keys) is ndarray.NDArray and type(grads) is ndarray.NDArray:
            assert grads.shape == (*keys.shape, self.width)
            wait = self.cache.embedding_update_raw(keys.data_ptr, grads.data_ptr,
                                                   int(np.prod(keys.shape)))
            wait._keep = (keys, grads)
        else:
            raise TypeError
        return self._done(wait, sync)

    def embedding_update_with_push_keys(self, keys, push_keys, grads, sync=False):
        if type(keys) is np.ndarray and type(grads) is np.ndarray:
            assert grads.shape == (keys.size, self.width)
            assert keys.dtype == np.uint64
            assert push_keys.dtype == np.uint64
            assert grads.dtype == np.float32
            wait = self.cache.embedding_update_with_push_keys(keys, push_keys, grads)
        elif type(keys) is ndarray.NDArray and type(grads) is ndarray.NDArray:
            assert grads.shape == (*keys.shape, self.width)
            n = int(np.prod(keys.shape))
            if isinstance(push_keys, np.ndarray):
                assert push_keys.dtype == np.uint64, \
                    "push_keys should be np.uint64, but currently is {}".format(push_keys.dtype)
                wait = self.cache.embedding_update_with_push_keys_np_raw(
                    keys.data_ptr, push_keys, grads.data_ptr, n)
            elif isinstance(push_keys, ndarray.NDArray):
                wait = self.cache.embedding_update_with_push_keys_raw(
                    keys.data_ptr, push_keys.data_ptr, grads.data_ptr, n,
                    int(np.prod(push_keys.shape)))
            else:
                raise TypeError
            wait._keep = (keys, push_keys, grads)
        else:
            raise TypeError
        return self._done(wait, sync)

    def embedding_push_pull(self, pullkeys, dest, pushkeys, grads, sync=False):
        if all(type(x) is ndarray.NDArray for x in (pullkeys, dest, pushkeys, grads)):
            assert grads.shape == (*pushkeys.shape, self.width)
            assert dest.shape == (*pullkeys.shape, self.width)
            wait = self.cache.embedding_push_pull_raw(
                pullkeys.data_ptr, dest.data_ptr, int(np.prod(pullkeys.shape)),
                pushkeys.data_ptr, grads.data_ptr, int(np.prod(pushkeys.shape)))
            wait._keep = (pullkeys, dest, pushkeys, grads)
        else:
            raise TypeError
        return self._done(wait, sync)

    @property
    def width(self):
        return self.cache.width

    @property
    def limit(self):
        return self.cache.limit

    def perf_enabled(self, enable=True):
        self.cache.perf_enabled = enable

    @property
    def perf(self):
        return self.cache.perf

    def flush(self):
        """Write every pending update back to the owner shard (checkpoint barrier)."""
        self.cache.flush()

    def bypass(self):
        self.cache.bypass()

    def undobypass(self):
        self.cache.undo_bypass()

    def __repr__(self):
        return self.cache.__repr__()

    # single-key debug calls
    def lookup(self, key):
        return self.cache.lookup(key)

    def count(self, key):
        return self.cache.count(key)

    def insert(self, embedding):
        return self.cache.insert(embedding)

    def keys(self):
        return self.cache.keys()

    def get_perf(self):
        return self.perf

    def _filtered(self, include_cold_start):
        return self.perf if include_cold_start else [x for x in self.perf if x["is_full"]]

    def overall_miss_rate(self, include_cold_start=False):
        """Miss rate of pulls (cstable.py:200-211)."""
        perf = self._filtered(include_cold_start)
        if not perf:
            return -1
        pulls = [x for x in perf if x["type"] == "Pull"]
        return np.sum([x["num_miss"] for x in pulls]) / np.sum([x["num_unique"] for x in pulls])

    def overall_data_rate(self, include_cold_start=False):
        """Rows transferred relative to a cache-less sparse pull/push (cstable.py:213-224)."""
        perf = self._filtered(include_cold_start)
        if not perf:
            return -1
        return np.sum([x["num_transfered"] for x in perf]) / np.sum([x["num_all"] for x in perf])
