"""Batch feeders either side of the embedding path (python/hetu/dataloader.py, python/hetu/laia/
laia_dataloader.py:172-259): batch order, the prefetch peek, data-parallel striding, the push index
of DataloaderWithPushIndex and the planner-driven LAIADataloader.  Host code, no GPU."""
import numpy as np

from herald_b200 import dataloader as dl
from herald_b200.laia import LAIAScheduler, TopkScheduler, laia_dataloader_op


class _Cfg(object):
    def __init__(self, rank=0, nrank=1, local_rank=0, limit=50, context_launch=False):
        self.rank, self.nrank, self.local_rank, self.cache_limit = rank, nrank, local_rank, limit
        self.context_launch = context_launch


def test_dataloader_batches_wrap_and_peek():
    data = np.arange(10 * 3, dtype=np.float32).reshape(10, 3)
    op = dl.dataloader_op([[data, 3, "train"]])
    op.backward_hook(_Cfg())
    assert op.get_batch_num("train") == 3 and op.get_cur_shape("train") == (3, 3)
    for step in range(8):
        b = step % 3
        nxt = op.get_next_arr("train").asnumpy()
        cur = op.get_arr("train").asnumpy()
        assert np.array_equal(nxt, cur)                      # the peek is the batch get_arr returns
        assert np.array_equal(cur, data[3 * b:3 * b + 3])    # drop_last: the 10th row is never served


def test_dataloader_data_parallel_stride():
    data = np.arange(40, dtype=np.float32).reshape(20, 2)
    for rank in range(2):
        op = dl.dataloader_op([dict(raw_data=data, batch_size=2, name="train")])
        op.backward_hook(_Cfg(rank=rank, nrank=2, context_launch=True))
        mine = data[rank::2]
        for b in range(4):
            assert np.array_equal(op.get_arr("train").asnumpy(), mine[2 * b:2 * b + 2])


def test_push_index_is_the_ascending_unique_ids_of_the_batch():
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 30, (12, 4)).astype(np.float32)
    op = dl.dataloader_with_push_index_op([[ids, 4, "train"]])
    op.backward_hook(_Cfg())
    for b in range(3):
        arr, push = op.get_arr("train")
        assert np.array_equal(arr.asnumpy(), ids[4 * b:4 * b + 4])
        assert push.dtype == np.uint64
        assert np.array_equal(push, np.unique(ids[4 * b:4 * b + 4]).astype(np.uint64))


def test_laia_dataloader_follows_the_planner():
    """Sparse ids, labels and dense features of batch b are the rows the planner assigned to this
    rank, and the sparse loader carries the plan of batch b + 1 (laia_dataloader.py:108-114,198-203)."""
    rng = np.random.default_rng(1)
    nrank, rank, batch, T = 2, 1, 4, 26
    S = nrank * batch * 10
    sparse = ((rng.zipf(1.2, (S, T)) - 1) % 150 + 1).astype(np.float32)
    labels = rng.integers(0, 2, (S, 1)).astype(np.float32)
    sched = LAIAScheduler(sparse, batch)
    ids_op = laia_dataloader_op([[sparse, batch, "train"]], sched, 0, is_sparse=True)
    y_op = laia_dataloader_op([[labels, batch, "train"]], sched, 1)
    sched.start(_Cfg(rank=rank, nrank=nrank, limit=40), dataset_num=2, epoch_num=1)
    for op in (ids_op, y_op):
        op.backward_hook(_Cfg(rank=rank, nrank=nrank))
    # what the bare planner computes for this rank
    from herald_b200.laia import LaiaScheduler
    ref = LaiaScheduler()
    emb = sparse.astype(np.intc).astype(np.int64)
    ref.start(emb, S, T, 1, sched.batch_size, sched.batch_num, nrank, rank, 40, 16)
    seq = []
    while ref.step():
        seq.append((ref.plan_of(rank).tolist(), ref.dist_of(rank).tolist()))
    for b in range(4):
        assert ids_op.get_cur_shape("train") == (sched.batch_size, T)
        ids, plan = ids_op.get_arr("train")
        y = y_op.get_arr("train")
        idx = seq[b][1]
        assert np.array_equal(ids.asnumpy(), sparse[idx])
        assert np.array_equal(y.asnumpy(), labels[idx])
        assert plan.dtype == np.uint64 and plan.tolist() == seq[b + 1][0]
