"""Batch feeders of the embedding path — the interface the cache ops and the Laia front end see
(python/hetu/dataloader.py:11-136,140-260): ``Dataloader`` (float32-carried data, one pinned
NDArray per queued batch), ``DataloaderWithPushIndex`` (a sparse batch plus the ids it touches, the
``push_keys`` of ``embedding_update_with_push_keys``), ``DataloaderOp`` (name -> loader dispatch:
``get_arr`` / ``get_next_arr`` / ``get_cur_shape`` / ``get_batch_num``) and the two factory
functions.  ``get_arr`` returns the current batch and advances; ``get_next_arr`` peeks at the batch
``get_arr`` will return next — what ParameterServerCommunicateOp prefetches
(gpu_ops/ParameterServerCommunicate.py:104-105).

Restated around a ring of pre-filled pinned buffers indexed by batch number modulo the ring size
(the reference re-maps a dict of slots); drop_last=True only, as the reference asserts."""
import numpy as np

from . import ndarray


class Dataloader(object):
    ring = 3                      # current batch, the prefetched next one, one being refilled

    def __init__(self, raw_data, batch_size, name="default", func=None, drop_last=True):
        assert drop_last, "drop_last must be True"          # dataloader.py:18
        self.func = func if func else (lambda x: x)
        self.raw_data = np.array(self.func(raw_data), np.float32)   # ids travel as float32 (:14)
        self.batch_size, self.drop_last, self.name = batch_size, drop_last, str(name)

    def init_states(self, rank=None, nrank=None):
        if rank is not None:      # data parallel: every nrank-th sample, starting at rank (:26)
            per = self.raw_data.shape[0] // nrank
            self.raw_data = self.raw_data[rank:per * nrank:nrank]
        self.samples_num = len(self.raw_data)
        self.queue_size = self.ring
        self.batch_size = min(int(self.batch_size), self.samples_num // self.ring)
        assert self.batch_size > 0, "Batch size %d invalid." % self.batch_size
        self.batch_num = self.samples_num // self.batch_size
        self.shape = (self.batch_size,) + tuple(self.raw_data.shape[1:])
        self.batch_index = 0      # batch get_arr() returns next
        self._filled = {}         # ring slot -> batch number it holds
        self._bufs = [ndarray.empty(self.shape, ctx=ndarray.cpu(0)) for _ in range(self.ring)]
        for b in range(self.ring - 1):
            self._fill(b)

    def _rows(self, batch):
        lo = (batch % self.batch_num) * self.batch_size
        return self.raw_data[lo:lo + self.batch_size]

    def _fill(self, batch):
        slot = batch % self.ring
        if self._filled.get(slot) != batch:
            self._bufs[slot][:] = self._rows(batch)
            self._filled[slot] = batch
        return self._bufs[slot]

    def _get_arr(self, batch):
        arr = self._fill(batch)
        self._fill(batch + 1)     # keep the next batch ready for get_next_arr
        return arr

    def get_arr(self):
        res = self._get_arr(self.batch_index)
        self.last_batch_size = self.batch_size
        self.batch_index += 1
        return res

    def get_next_arr(self):
        return self._get_arr(self.batch_index)

    def get_cur_shape(self):
        return self.shape


class DataloaderWithPushIndex(Dataloader):
    """(batch, ids of the batch as uint64) — dataloader.py:200-247.  The ids come out ascending and
    unique, which is what the cache's plan merge expects (cache.cc:286-301)."""

    def _get_arr(self, batch):
        arr = super()._get_arr(batch)
        return arr, np.unique(self._rows(batch).reshape(-1)).astype(np.uint64)

    def get_cur_shape(self):
        return self.shape


class DataloaderOp(object):
    """Graph node that owns one loader per data set name (dataloader.py:140-182)."""

    def __init__(self, dataloaders):
        self.dataloaders = {dl.name: dl for dl in dataloaders}
        self.name = "DataloaderOp(%s)" % "_".join(self.dataloaders)
        self.inputs, self.ctx = [], ndarray.cpu(0)
        self.on_gpu, self.on_cpu = False, True

    def get_batch_num(self, name):
        return self.dataloaders[name].batch_num

    def get_arr(self, name):
        return self.dataloaders[name].get_arr()

    def get_next_arr(self, name):
        return self.dataloaders[name].get_next_arr()

    def get_cur_shape(self, name):
        return self.dataloaders[name].get_cur_shape()

    def forward_hook(self, config):
        pass

    def backward_hook(self, config):
        for d in self.dataloaders.values():
            if getattr(config, "context_launch", False):
                d.init_states(config.rank, config.nrank)
            else:
                d.init_states()


def _build(cls, spec):
    if isinstance(spec, Dataloader):
        return spec
    if isinstance(spec, (list, tuple)):
        return cls(*spec)
    if isinstance(spec, dict):
        return cls(**spec)
    raise AssertionError("Dataloader parameter invalid.")


def dataloader_op(dataloaders):
    return DataloaderOp([_build(Dataloader, d) for d in dataloaders])


def dataloader_with_push_index_op(dataloaders):
    return DataloaderOp([_build(DataloaderWithPushIndex, d) for d in dataloaders])
