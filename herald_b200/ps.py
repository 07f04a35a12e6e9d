"""Worker-side "communicate" object — the slice of ps-lite's Python binding that the embedding
hot path touches (ps-lite/src/python_binding.cc:94-132, used at python/hetu/cstable.py:21,36 and
python/hetu/initializers.py:28-38).  The parameter server of the reference becomes a table shard
in this process's HBM (hb_table_*); BarrierWorker becomes an NCCL barrier when a group is up.
"""
import ctypes

import numpy as np

from ._base import _LIB, check_call

_sz = ctypes.c_size_t
_vp = ctypes.c_void_p

# ps::ParamType / ps::InitType (ps-lite/include/ps/server/param.h:11-15, psf/misc.h:7-12)
kParam, kParam2D, kCacheTable = 0, 1, 2
Constant, Uniform, Normal, TruncatedNormal = 0, 1, 2, 3

_LIB.hb_table_create.argtypes = [ctypes.c_int, _sz, _sz, ctypes.c_int, ctypes.POINTER(_vp)]
_LIB.hb_table_get.argtypes = [ctypes.c_int, ctypes.POINTER(_vp)]
_LIB.hb_table_destroy.argtypes = [_vp]
_LIB.hb_table_init.argtypes = [_vp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                               ctypes.c_ulonglong]
_LIB.hb_table_load_rows.argtypes = [_vp, _sz, _sz, _vp]
_LIB.hb_table_read_rows.argtypes = [_vp, _sz, _sz, _vp]
_LIB.hb_table_read_versions.argtypes = [_vp, _sz, _sz, _vp]
_LIB.hb_table_read_rows_at.argtypes = [_vp, _vp, _sz, _vp, _vp]
_LIB.hb_table_shard.argtypes = [_vp, ctypes.POINTER(_sz), ctypes.POINTER(_sz),
                                ctypes.POINTER(_vp), ctypes.POINTER(_vp)]
_LIB.hb_table_save.argtypes = [_vp, ctypes.c_char_p]
_LIB.hb_table_load.argtypes = [_vp, ctypes.c_char_p]
_LIB.hb_comm_rank.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]


class Table(object):
    """Handle of one registered table (this rank's shard of it)."""

    def __init__(self, handle, node_id, length, width):
        self.h, self.node_id, self.length, self.width = handle, node_id, length, width

    def shard(self):
        begin, rows, dr, dv = _sz(), _sz(), _vp(), _vp()
        check_call(_LIB.hb_table_shard(self.h, ctypes.byref(begin), ctypes.byref(rows),
                                       ctypes.byref(dr), ctypes.byref(dv)))
        return begin.value, rows.value, dr.value, dv.value

    def load_rows(self, rows, row_begin=0):
        rows = np.ascontiguousarray(rows, np.float32)
        assert rows.ndim == 2 and rows.shape[1] == self.width
        check_call(_LIB.hb_table_load_rows(self.h, row_begin, rows.shape[0], rows.ctypes.data))

    def read_rows(self, row_begin=None, nrows=None):
        """Rows of the local shard by default (global range when given)."""
        sb, sn, _, _ = self.shard()
        row_begin = sb if row_begin is None else row_begin
        nrows = sn if nrows is None else nrows
        out = np.zeros((nrows, self.width), np.float32)
        check_call(_LIB.hb_table_read_rows(self.h, row_begin, nrows, out.ctypes.data))
        return out

    def read_rows_at(self, keys):
        """(rows, versions) of the given global rows — sparse pull from this rank's shard."""
        keys = np.ascontiguousarray(keys, np.uint64).reshape(-1)
        rows = np.zeros((keys.size, self.width), np.float32)
        ver = np.zeros(keys.size, np.int64)
        check_call(_LIB.hb_table_read_rows_at(self.h, keys.ctypes.data, keys.size, rows.ctypes.data,
                                              ver.ctypes.data))
        return rows, ver

    def read_versions(self, row_begin=None, nrows=None):
        sb, sn, _, _ = self.shard()
        row_begin = sb if row_begin is None else row_begin
        nrows = sn if nrows is None else nrows
        out = np.zeros(nrows, np.int64)
        check_call(_LIB.hb_table_read_versions(self.h, row_begin, nrows, out.ctypes.data))
        return out


class WorkerCommunicate(object):
    """Names follow the reference binding so call sites read the same."""

    def __init__(self, device=0):
        self.device = device
        self.tables = {}

    def rank(self):
        r, w = ctypes.c_int(), ctypes.c_int()
        _LIB.hb_comm_rank(ctypes.byref(r), ctypes.byref(w))
        return r.value

    def nrank(self):
        r, w = ctypes.c_int(), ctypes.c_int()
        _LIB.hb_comm_rank(ctypes.byref(r), ctypes.byref(w))
        return w.value

    def InitTensor(self, node_id, ptype, length, width, init_type=Constant, init_a=0.0,
                   init_b=1.0, seed=0, opt_type=0, opt_args=None):
        """InitTensor(id, kCacheTable, len, width, init_type, a, b, seed, ...)
        (python/hetu/initializers.py:28-38).  Only sparse/cache tables are on this path."""
        if ptype not in (kParam2D, kCacheTable):
            raise NotImplementedError("dense parameters are outside the embedding hot path")
        h = _vp()
        check_call(_LIB.hb_table_create(int(node_id), int(length), int(width), self.device,
                                        ctypes.byref(h)))
        check_call(_LIB.hb_table_init(h, int(init_type), float(init_a), float(init_b), int(seed)))
        t = Table(h, int(node_id), int(length), int(width))
        self.tables[int(node_id)] = t
        return t

    def table(self, node_id):
        return self.tables[int(node_id)]

    def ClearTensor(self, node_id):
        t = self.tables.pop(int(node_id), None)
        if t is not None:
            check_call(_LIB.hb_table_destroy(t.h))

    def SaveParam(self, node_id, address):
        """SaveParam(node, dir) (ps-lite/src/python_binding.cc:111-113): every shard writes
        "<dir>/<node>_<partition>.dat", raw row-major float32 (PSAgent.h:447-460)."""
        check_call(_LIB.hb_table_save(self.tables[int(node_id)].h, str(address).encode()))

    def LoadParam(self, node_id, address):
        """LoadParam(node, dir) (ps-lite/src/python_binding.cc:115-117, PSAgent.h:462-476)."""
        check_call(_LIB.hb_table_load(self.tables[int(node_id)].h, str(address).encode()))

    def BarrierWorker(self):
        if self.nrank() > 1:
            check_call(_LIB.hb_comm_barrier())

    def Wait(self, node_id):
        pass


_LIB.hb_comm_unique_id.argtypes = [ctypes.c_char_p]
_LIB.hb_comm_init.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]


def group_init(rank, world, device, exchange):
    """Bring up the multi-GPU group (one process per GPU of one NVSwitch box).  `exchange(b)`
    broadcasts rank 0's 128 bytes to every rank — any rendezvous the launcher has: the bench and
    the tests use torch.distributed (gloo) for it; nothing else of torch is involved.  Tables and
    caches created afterwards are sharded / exchange over NVLink; their create / reserve /
    update / destroy calls are collective (same order on every rank)."""
    idbuf = ctypes.create_string_buffer(128)
    if rank == 0:
        check_call(_LIB.hb_comm_unique_id(idbuf))
    uid = exchange(bytes(idbuf.raw))
    check_call(_LIB.hb_comm_init(uid, int(rank), int(world), int(device)))


def group_finalize():
    check_call(_LIB.hb_comm_finalize())


_comm = None


def worker_init(device=0):
    global _comm
    if _comm is None:
        _comm = WorkerCommunicate(device)
    return _comm


def get_worker_communicate():
    """hetu.get_worker_communicate(): the process-wide communicate object."""
    return worker_init()


def worker_finish():
    global _comm
    if _comm is not None:
        for nid in list(_comm.tables):
            _comm.ClearTensor(nid)
        _comm = None
