"""Multi-GPU parity: N ranks, row-sharded table, pushes through owner mailboxes over
NVLink, rows pulled straight from the owners' shards — against a whole-group replay on the oracle
(tests/mg_worker.py).  Needs >= 2 GPUs on the box; `gpurun --gpus 2 -- python -m pytest
tests/test_multi_gpu.py -m gpu` is how it is run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import ctypes
    n = ctypes.c_int(0)
    try:
        ctypes.CDLL("libcudart.so.12").cudaGetDeviceCount(ctypes.byref(n))
    except OSError:
        return 0
    return n.value


def _run(world, **env):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    port = 29600 + (os.getpid() + hash(tuple(sorted(env.items())))) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mg_worker.py")]
    r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok (") == world, r.stdout[-2000:]


@pytest.mark.parametrize("policy,bound", [("lru", 0), ("lru", 3), ("lfu", 0), ("lfuopt", 2)])
def test_group_of_two(policy, bound):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, MG_POLICY=policy, MG_BOUND=bound)


def test_two_tables_interleaved():
    """Two caches per rank, their exchanges enqueued back to back (per-cache exchange flags)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, MG_POLICY="lru", MG_BOUND=0, MG_TABLES=2, MG_STEPS=15)


def test_bench_shape_rows_through_the_mailbox():
    """D = 128 rows and thousands of keys per call through the owner mailboxes and remote pulls."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, MG_POLICY="lru", MG_BOUND=0, MG_V=50021, MG_D=128, MG_LIMIT=6000, MG_N=20000, MG_STEPS=8)


def test_group_of_all_gpus():
    n = min(_ngpu(), 8)
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    _run(n, MG_POLICY="lru", MG_BOUND=0, MG_V=4099, MG_LIMIT=300, MG_N=600)
