"""herald_b200 — B200-native embedding hot path behind Hetu/Herald's operator and cache API.

Python here is host glue only: ctypes calls into ``lib/libherald_b200.so`` (hand-written CUDA
for sm_100a).  There is no CPU fallback: importing any compute module without the built
library, or calling it without a GPU, raises.

Reference surface mirrored (paths relative to the reference root):
  python/hetu/ndarray.py      -> herald_b200.ndarray   (NDArray, IndexedSlices, cpu/gpu/array/empty)
  python/hetu/stream.py       -> herald_b200.stream    (Stream, Event, CSEvent)
  python/hetu/gpu_links/*     -> herald_b200.gpu_links (embedding_lookup, sgd_update, adam_update, ...)
  build/lib/hetu_cache*.so    -> herald_b200.hetu_cache (LRUCache/LFUCache/LFUOptCache, Embedding)
  python/hetu/cstable.py      -> herald_b200.cstable   (CacheSparseTable)
  ps worker communicate       -> herald_b200.ps        (InitTensor / BarrierWorker on HBM shards)
  python/hetu/gpu_ops/EmbeddingLookUp.py -> herald_b200.gpu_ops (embedding_lookup_op)
"""
from . import ndarray
from .ndarray import cpu, gpu, array, empty, is_gpu_ctx, NDArray, IndexedSlices
from .ps import get_worker_communicate, worker_init, worker_finish
from .gpu_ops import embedding_lookup_op, embedding_lookup_gradient_op

__all__ = ["ndarray", "cpu", "gpu", "array", "empty", "is_gpu_ctx", "NDArray", "IndexedSlices",
           "get_worker_communicate", "worker_init", "worker_finish", "embedding_lookup_op",
           "embedding_lookup_gradient_op"]
