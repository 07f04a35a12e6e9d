/*
 * herald_b200.h — C-ABI of libherald_b200.so, the B200 (sm_100a) implementation of
 * Herald/Hetu's data-parallel embedding hot path.
 *
 * Plain C: pointers and sizes only.  Two groups of entry points:
 *
 *  (b1) the symbols Hetu's link layer binds from libc_runtime_api.so through
 *       ctypes (python/hetu/_base.py:66-77), with the reference's names, argument
 *       order and DLArray/DLStream structs — a drop-in for those symbols;
 *  (b2) hb_* functions: the owner-side table shard and the worker-side cache, which
 *       replace the pybind module `hetu_cache` (src/hetu_cache/src/python_api.cc)
 *       and the three ps-lite transport calls below it
 *       (ps-lite/include/ps/worker/hetu_binding.h:14-28).  The Python module
 *       herald_b200.hetu_cache presents the reference's class surface on top.
 *
 * Conventions: every function returns 0 on success and -1 on failure;
 * HBGetLastError() returns the message of the calling thread's last failure
 * (the reference returns -1 from API_BEGIN/API_END, src/common/runtime_base.h:13-47).
 * All device work is asynchronous with respect to the host and ordered on the
 * stream named in the call (NULL = the legacy default stream for b1, the cache's
 * own stream for b2).  All citations are relative to the reference root.
 */
#ifndef HERALD_B200_H_
#define HERALD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* DLArray ABI — replaces src/common/dlarray.h:22-65                           */
/* ------------------------------------------------------------------------- */
typedef enum { kCPU = 1, kGPU = 2 } DLDeviceType;
typedef struct {
    int device_id;
    DLDeviceType device_type;
} DLContext;
typedef struct {
    void *data;
    DLContext ctx;
    int ndim;
    int64_t *shape;
    int64_t *stride;
} DLArray;
typedef struct {
    int device_id;
    void *handle; /* points to a cudaStream_t (src/ops/EmbeddingLookup.cu:46) */
} DLStream;
typedef struct {
    int device_id;
    void *handle; /* points to a cudaEvent_t */
} DLEvent;
typedef int64_t index_t;
typedef DLArray *DLArrayHandle;
typedef DLStream *DLStreamHandle;
typedef DLEvent *DLEventHandle;

const char *HBGetLastError(void);
/* Library/arch identification: "herald_b200 <ver> sm_100a". */
const char *HBVersion(void);
/* Number of kernels this library has launched in this process (bench `gpu_launches`). */
uint64_t HBKernelLaunchCount(void);
/* Segments (occurrences of one id in a batch) longer than `rows` are reduced by the column-split
 * hot path; default 64 (or $HERALD_HOT_THRESHOLD).  Applies to op-level calls and to caches
 * created afterwards.  Results are bit-identical either way; tests lower it for coverage. */
int HBSetHotThreshold(unsigned rows);
/* Diagnostics: per-CTA / per-hot-row timeline (globaltimer ns) of the most recent segment-reduce
 * launch.  Layout: [0] grid, [1] hot items, 4 words per CTA {start, hot phase end, end, items},
 * then 3 words per hot item {start, end, occurrences}.  Off by default (no cost when off). */
int HBSegTraceEnable(int on);
int HBSegTraceRead(unsigned long long *out, size_t words);

/* ---- runtime plumbing: src/common/c_runtime_api.h:28-77 ------------------- */
int DLStreamCreate(size_t dev_id, DLStreamHandle *handle);
int DLStreamDestroy(DLStreamHandle handle);
int DLStreamSync(DLStreamHandle handle);
int DLEventCreate(size_t dev_id, DLEventHandle *handle);
int DLEventDestroy(DLEventHandle handle);
int DLEventRecord(DLStreamHandle stream_handle, DLEventHandle event_handle);
int DLEventSync(DLEventHandle handle);
int DLEventElapsedTime(DLEventHandle start, DLEventHandle ending, float *duration);
/* CPU arrays are allocated in pinned host memory so that H2D/D2H copies are async. */
int DLArrayAlloc(const index_t *shape, const index_t *stride, index_t ndim, DLContext ctx,
                 DLArrayHandle *out);
int DLArrayFree(DLArrayHandle handle);
int DLArrayCopyFromTo(DLArrayHandle from, DLArrayHandle to, DLStreamHandle stream);
int DLGpuArraySet(DLArrayHandle arr, float value, DLStreamHandle stream_handle);

/* ---- op-level hot path (table resident in HBM; ids are float32 arrays) ---- */
/* src/common/c_runtime_api.h:308-310; kernel src/ops/EmbeddingLookup.cu:3-52 */
int DLGpuEmbeddingLookUp(const DLArrayHandle input, const DLArrayHandle ids,
                         DLArrayHandle output, DLStreamHandle stream_handle);
/* c_runtime_api.h:312-314; src/ops/EmbeddingLookup.cu:54-131 (zero + scatter-add) */
int DLGpuEmbeddingLookUp_Gradient(const DLArrayHandle output_grad, const DLArrayHandle ids,
                                  DLArrayHandle input_grad, DLStreamHandle stream_handle);
/* c_runtime_api.h:700-702; src/ops/OptimizersSparse.cu:282-329 */
int DeduplicateIndexedSlices(const DLArrayHandle origin, const DLArrayHandle inverse,
                             DLArrayHandle compressed, DLStreamHandle stream_handle);
/* c_runtime_api.h:704-706; src/ops/OptimizersSparse.cu:233-280 */
int IndexedSlices2Dense(const DLArrayHandle values, const DLArrayHandle indices,
                        DLArrayHandle new_values, DLStreamHandle stream_handle);
/* c_runtime_api.h:569-571; src/ops/IndexedSlices.cu:3-48 */
int IndexedSlicesOneSideAdd(const DLArrayHandle indices, const DLArrayHandle values,
                            DLArrayHandle output, DLStreamHandle stream_handle);
/* c_runtime_api.h:639-642; src/ops/OptimizersSparse.cu:3-51 */
int AddL2RegularizationSparse(const DLArrayHandle param, const DLArrayHandle grad_indices,
                              DLArrayHandle grad_values, float l2reg,
                              DLStreamHandle stream_handle);
/* c_runtime_api.h:645-648; src/ops/OptimizersSparse.cu:53-99 (duplicate ids allowed) */
int SGDOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                             const DLArrayHandle grad_values, float lr,
                             DLStreamHandle stream_handle);
/* c_runtime_api.h:653-656; src/ops/OptimizersSparse.cu:101-229 */
int MomentumOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                                  const DLArrayHandle grad_values, DLArrayHandle velocity,
                                  float lr, float momentum, bool nesterov,
                                  DLStreamHandle stream_handle);
/* c_runtime_api.h:661-664; src/ops/OptimizersSparse.cu:331-389 (ids unique) */
int AdaGradOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                                 const DLArrayHandle grad_values, DLArrayHandle acc, float lr,
                                 float eps, DLStreamHandle stream_handle);
/* c_runtime_api.h:670-674; src/ops/OptimizersSparse.cu:391-455 (ids unique) */
int AdamOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                              const DLArrayHandle grad_values, DLArrayHandle expavg,
                              DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                              float beta1t, float beta2t, float eps,
                              DLStreamHandle stream_handle);
/* c_runtime_api.h:681-686; src/ops/OptimizersSparse.cu:457-522 (ids unique) */
int AdamWOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                               const DLArrayHandle grad_values, DLArrayHandle expavg,
                               DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                               float beta1t, float beta2t, float eps, float weight_decay,
                               DLStreamHandle stream_handle);
/* c_runtime_api.h:692-698, src/ops/OptimizersSparse.cu:596-721: Adam direction scaled by the trust
 * ratio ||param[ids]|| / ||update|| (norms over the listed rows only); ids must be unique. */
int LambOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                               const DLArrayHandle grad_values, DLArrayHandle expavg,
                               DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                               float beta1t, float beta2t, float eps, float weight_decay,
                               DLStreamHandle stream_handle);

/* ---- device-side dedup: replaces the host np.unique round trip of
 *      IndexedSlices.deduplicate (python/hetu/ndarray.py:532-554) ----------- */
/* ids: float32[n] on the GPU.  unique_ids: float32[>=n] (ascending), inverse: float32[n]
 * (rank of ids[i] in unique_ids, carried as float32 like the reference), num_unique: one
 * int64 in DEVICE memory.  Everything stays on `stream_handle`; no host sync. */
int HBUniqueIndexedSlices(const DLArrayHandle ids, DLArrayHandle unique_ids,
                          DLArrayHandle inverse, int64_t *num_unique_dev,
                          DLStreamHandle stream_handle);
/* Fused np.unique + DeduplicateIndexedSlices + AdamOptimizerSparseUpdate on ids with
 * duplicates: one sort, one deterministic segment reduce, one row update. */
int HBAdamSparseUpdateFused(DLArrayHandle param, const DLArrayHandle grad_indices,
                            const DLArrayHandle grad_values, DLArrayHandle expavg,
                            DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                            float beta1t, float beta2t, float eps,
                            DLStreamHandle stream_handle);

/* ------------------------------------------------------------------------- */
/* (b2) owner-side table shard — replaces the PS server's CacheTable           */
/*      ps-lite/include/ps/server/param.h:119-138 and the handlers              */
/*      ps-lite/src/PSFhandle_embedding.cc:5-79                                 */
/* ------------------------------------------------------------------------- */
typedef struct hb_table hb_table;
typedef struct hb_cache hb_cache;

/* Register table `node_id` (the server key, as InitTensor: ps-lite/src/python_binding.cc:94)
 * of `length` x `width` fp32 rows on `device`.  In a multi-GPU group (hb_comm_init called
 * first) each rank allocates only its AveragePartitioner row range
 * (ps-lite/include/ps/partitioner.h:46-57).  Rows and versions start at zero. */
int hb_table_create(int node_id, size_t length, size_t width, int device, hb_table **out);
int hb_table_get(int node_id, hb_table **out);
int hb_table_destroy(hb_table *t);
/* init_type as ps::InitType (ps-lite/include/ps/psf/misc.h:7-12): 0 constant(a),
 * 1 uniform[a,b), 2 normal(a,b), 3 truncated normal(a,b) within 2 sigma.  Counter-based
 * generator keyed on (seed, global row, column): independent of the sharding. */
int hb_table_init(hb_table *t, int init_type, double a, double b, unsigned long long seed);
/* Copy rows [row_begin, row_begin+nrows) of the GLOBAL table from/to host or device memory;
 * only the part owned by this rank is touched.  Synchronous. */
int hb_table_load_rows(hb_table *t, size_t row_begin, size_t nrows, const float *rows);
int hb_table_read_rows(hb_table *t, size_t row_begin, size_t nrows, float *rows);
int hb_table_read_versions(hb_table *t, size_t row_begin, size_t nrows, int64_t *versions);
/* Sparse read — the owner side of ps-lite's SparsePull (ps-lite/include/ps/psf/sparse.h:9-20, ps-lite/include/ps/worker/PSAgent.h:205-230) without
 * a cache in front: rows[i,:] / versions[i] of GLOBAL row keys[i], n keys in host or device
 * memory, results to host or device memory (either may be NULL).  Keys outside this rank's shard
 * yield zeros / version -1.  Synchronous. */
int hb_table_read_rows_at(hb_table *t, const uint64_t *keys, size_t n, float *rows, int64_t *versions);
/* local shard geometry */
/* Checkpoint of this rank's shard in the reference's on-disk format
 * (ps-lite/include/ps/worker/PSAgent.h:447-476 ParameterSave/ParameterLoad,
 * ps-lite/include/ps/server/PSFHandle.h:401-439): "<dir>/<node_id>_<partition>.dat", raw row-major
 * float32, partition = rank.  "<dir>/<node_id>_<partition>.ver" (int64 row versions) is an
 * extension the loader treats as optional.  Collective in a group (every rank its own file). */
int hb_table_save(hb_table *t, const char *dir);
int hb_table_load(hb_table *t, const char *dir);
int hb_table_shard(hb_table *t, size_t *row_begin, size_t *nrows, float **dev_rows,
                   int64_t **dev_versions);

/* ------------------------------------------------------------------------- */
/* (b2) worker-side cache — replaces hetu_cache LRUCache/LFUCache/LFUOptCache  */
/*      src/hetu_cache/src/python_api.cc:32-76, src/hetu_cache/src/cache.cc     */
/* ------------------------------------------------------------------------- */
enum { HB_POLICY_LRU = 0, HB_POLICY_LFU = 1, HB_POLICY_LFUOPT = 2 };
/* key encodings accepted by the batch entry points */
enum {
    HB_KEYS_U64 = 0, /* uint64 keys: the numpy path (python/hetu/cstable.py:51-55) */
    HB_KEYS_F32 = 1  /* float32-carried ids, key = (uint64)(float)id: the NDArray "raw"
                        path (src/hetu_cache/src/cache.cc:49-58) */
};

/* counters of one call — the hit/miss parity surface (cache.cc:89-106, :179-196) */
typedef struct {
    int64_t num_all;        /* keys in the call                                     */
    int64_t num_unique;     /* distinct keys                                        */
    int64_t num_miss;       /* unique keys not found by the policy lookup           */
    int64_t num_evict;      /* dirty victims flushed by this push (Push only)       */
    int64_t num_transfered; /* Pull: rows the owner returned; Push: lines pushed    */
    int64_t is_full;        /* size() == limit after the call                       */
    int64_t size;           /* lines resident after the call                        */
    int64_t error;          /* non-zero: device-side failure code (see hb_cache.cu) */
    float time_ms;          /* device time of the call (CUDA events)                */
    float sort_ms, lookup_ms, transfer_ms, copy_ms, insert_ms; /* phase split      */
    float kernel_ms;        /* Push: the accumulate kernel alone (copy_ms minus its plan
                               kernel); 0 when the call carried no phase events        */
    int64_t num_remote;     /* multi-GPU: rows pulled from / lines pushed to a PEER GPU   */
} hb_perf;

/* limit/len/width/node_id as the reference constructors (python_api.cc:54-76).
 * The table `node_id` must exist.  pull_bound = push_bound = 5 initially (cache.h:26-27). */
int hb_cache_create(int policy, size_t limit, size_t length, size_t width, int node_id,
                    hb_cache **out);
int hb_cache_destroy(hb_cache *c);
int hb_cache_set_bounds(hb_cache *c, int64_t pull_bound, int64_t push_bound);
int hb_cache_get_bounds(hb_cache *c, int64_t *pull_bound, int64_t *push_bound);
int hb_cache_set_bypass(hb_cache *c, int on);                   /* cache.cc:15-35 */
/* Fold of -lr into the update (python/hetu/gpu_ops/ParameterServerCommunicate.py:24,58-59: the
 * reference multiplies the whole sparse gradient by -lr ON THE HOST before every push).  With a
 * scale set, hb_cache_update* / push_pull take the RAW gradient and every value is multiplied by
 * (float)scale inside the accumulate kernel — one exact fp32 product, then the exact add: bit-
 * identical to scaling first.  Default 1 (gradients arrive scaled). */
int hb_cache_set_grad_scale(hb_cache *c, float scale);
int hb_cache_get_grad_scale(hb_cache *c, float *scale);
/* Order in which one row's gradient occurrences are added by hb_cache_update*:
 *   0 (default)  occurrence order, ((row + g0) + g1) + ... — the reference's loop
 *                (src/hetu_cache/include/embedding.h:78-91): updated rows are BIT-IDENTICAL to it;
 *   1            rows with more than 1024 occurrences in the call are summed in 8 fixed runs of
 *                consecutive occurrences (each in occurrence order, from 0), the run sums added to
 *                the row in run order.  Still deterministic — the shape depends on the occurrence
 *                count only — but re-associated: those rows agree with the reference within 1e-5
 *                relative (BASELINE north star), all others stay bit-identical.  At WDL-Criteo
 *                batch 8192 one id occurs 11 000 times; in mode 0 that one dependent FADD chain,
 *                not HBM, is what bounds the update kernel. */
int hb_cache_set_reduce_mode(hb_cache *c, int mode);
int hb_cache_get_reduce_mode(hb_cache *c, int *mode);
/* Device-pointer callers: order the cache's streams behind `producer_stream` (a cudaStream_t),
 * i.e. behind the kernels that are still writing the keys / gradients of the next call — what the
 * executor's stream_handle + event is for in ParameterServerCommunicate.py:27-35. */
int hb_cache_after_stream(hb_cache *c, void *producer_stream);
/* perf_enabled (python_api.cc:40-41): also records the per-phase CUDA events behind
 * hb_perf.sort_ms / lookup_ms / transfer_ms / copy_ms / insert_ms. */
int hb_cache_set_perf(hb_cache *c, int on);
/* Phase events (sort / lookup / transfer / copy / insert times) only on every `every`-th pair of
 * calls: an event between two kernels costs their launch overlap.  Counters are always recorded;
 * an unsampled call reports zero phase times.  Default 1 (every call, as the reference). */
int hb_cache_set_perf_sampling(hb_cache *c, unsigned every);
/* Reserve workspace for calls of up to max_keys keys (grows on demand otherwise). */
int hb_cache_reserve(hb_cache *c, size_t max_keys);
/* The CUDA stream (cudaStream_t) all of this cache's work is ordered on. */
int hb_cache_stream(hb_cache *c, void **stream);

/* embedding_lookup[_raw]: cache.cc:37-107.  keys: n keys (host or device memory, encoding
 * `key_kind`); dest: n*width floats (host or device).  Asynchronous: returns after
 * enqueueing; hb_cache_wait() completes it and fills `perf` (may be NULL). Caller keeps
 * keys/dest alive until then (python/hetu/cstable.py:38-45). */
int hb_cache_lookup(hb_cache *c, const void *keys, int key_kind, size_t n, float *dest);
/* embedding_update[_raw]: cache.cc:109-196.  grads: n*width floats, already scaled by -lr
 * (python/hetu/gpu_ops/ParameterServerCommunicate.py:24,58-59). */
int hb_cache_update(hb_cache *c, const void *keys, int key_kind, size_t n, const float *grads);
/* embedding_update_with_push_keys[_np_raw|_raw]: cache.cc:198-334.  push_keys ascending. */
int hb_cache_update_with_push_keys(hb_cache *c, const void *keys, int key_kind, size_t n,
                                   const void *push_keys, int push_key_kind, size_t n_push,
                                   const float *grads);
/* embedding_push_pull_raw: cache.cc:336-422 */
int hb_cache_push_pull(hb_cache *c, const void *pull_keys, int pull_kind, size_t n_pull,
                       float *dest, const void *push_keys, int push_kind, size_t n_push,
                       const float *grads);
/* Block until every call enqueued so far has completed; *perf receives the counters of the
 * most recent call.  Returns -1 if any of them failed on the device (each failure is reported
 * once; the cache stays usable). */
int hb_cache_wait(hb_cache *c, hb_perf *perf);
/* Per-call completion — what the reference's wait_t of ONE call is (python_api.cc:16-19,
 * cache.h:14).  hb_cache_last_call: sequence number of the call enqueued last.
 * hb_cache_wait_call: block until call `seq` (and, for a lookup into host memory, the download of
 * its rows) has completed; later calls keep running.  A host caller can so overlap the upload of
 * update(t+1)'s gradients with the download of lookup(t+1)'s rows (separate copy streams). */
int hb_cache_last_call(hb_cache *c, uint64_t *seq);
int hb_cache_wait_call(hb_cache *c, uint64_t seq, hb_perf *perf);
/* Counters of calls [first, first + count) (waits for the last of them), oldest first. */
int hb_cache_perf_range(hb_cache *c, uint64_t first, int count, hb_perf *out, int *kinds);
/* Push every dirty line (pending victims and resident lines with updates != 0) to its owner,
 * whatever the bound, and mark it clean: call before hb_table_save so that the checkpoint holds
 * every update (the reference lacks this: SURVEY section 5).  Synchronous; collective in a group. */
int hb_cache_flush(hb_cache *c);
/* Counters of the last `max` completed calls, oldest first; returns how many were written. */
int hb_cache_perf_history(hb_cache *c, hb_perf *out, int *kinds, int max, int *written);

/* Laia / Herald scoring against the REAL cache.  The reference's planners score a
 * sample for a worker by counting its embedding ids in a host-side simulation of that worker's cache
 * (MiniLRUCache snapshots: laia/src/laia_scheduler.cc:171-210, topk_scheduler.cc:405-428); here the
 * worker answers from its index in HBM.  sample_ids: [num_samples, num_tables] ids (host or device,
 * encoding key_kind); scores[i] = how many of sample i's tables table_order[0..top_k) (NULL = tables
 * 0..top_k-1) hold an id that has a line in this cache — with fresh != 0 only lines a lookup would
 * not re-pull (owner version - line version <= pull_bound), the snapshots' "valid" bit.  Read-only;
 * ordered on the cache's stream; with a host `scores` pointer the call completes before returning.
 * Workers exchange their score columns (an all-gather of num_samples words) and run the greedy
 * assignment on identical inputs. */
int hb_cache_score(hb_cache *c, const void *sample_ids, int key_kind, size_t num_samples, size_t num_tables,
                   const uint32_t *table_order, size_t top_k, int fresh, uint32_t *scores);
/* resident[i] = 1 when keys[i] has a line in the cache (CacheBase::count for a batch of keys): the
 * plan rule "ids of my samples that I hold" (topk_scheduler.cc:476-483) against the real index. */
int hb_cache_probe(hb_cache *c, const void *keys, int key_kind, size_t n, uint8_t *resident);

/* debug surface: python_api.cc:56-60 */
int hb_cache_size(hb_cache *c, size_t *size);
int hb_cache_count(hb_cache *c, uint64_t key, int *count);
int hb_cache_keys(hb_cache *c, uint64_t *keys, size_t capacity, size_t *n); /* ascending */
/* read one line without touching replacement state; *found = 0 if absent */
int hb_cache_peek(hb_cache *c, uint64_t key, int *found, int64_t *version, int64_t *updates,
                  float *data, float *grad);
/* single-key policy lookup (touches replacement state like CacheBase::lookup) */
int hb_cache_touch(hb_cache *c, uint64_t key, int *found, int64_t *version, float *data);
/* insert(Embedding): policy insert of one line with given version and data */
int hb_cache_insert(hb_cache *c, uint64_t key, int64_t version, const float *data);

/* ------------------------------------------------------------------------- */
/* multi-GPU: one process per GPU; replaces ps-lite's worker<->server transport */
/* (PSAgent row-range split ps-lite/include/ps/worker/PSAgent.h:537-627) with   */
/* NCCL grouped send/recv over NVLink.                                          */
/* ------------------------------------------------------------------------- */
/* 128-byte ncclUniqueId made on rank 0 and distributed by the launcher. */
int hb_comm_unique_id(void *id128);
int hb_comm_init(const void *id128, int rank, int world, int device);
int hb_comm_rank(int *rank, int *world);
int hb_comm_barrier(void); /* BarrierWorker (python/hetu/cstable.py:36) */
int hb_comm_finalize(void);

/* ------------------------------------------------------------------------- */
/* Laia / Herald embedding scheduler (host side) — replaces the pybind module    */
/* `laia_cache` (laia/src/python_binding.cc:8-16: LaiaScheduler.start / pop /    */
/* length) and the Cython planner python/hetu/laia/laia.pyx.  Every worker runs  */
/* the same planner; the plan of a worker is the `push_keys` argument of         */
/* hb_cache_update_with_push_keys.                                               */
/* ------------------------------------------------------------------------- */
typedef struct hb_laia hb_laia;
/* LaiaScheduler::start (laia/src/laia_scheduler.cc:30-88): sample_embs is the [num_sample,
 * num_table] matrix of embedding ids, copied.  A global batch is mini_batch_size * nrank samples. */
int hb_laia_create(hb_laia **out, const uint64_t *sample_embs, size_t num_sample, size_t num_table,
                   size_t epoch_num, size_t mini_batch_size, size_t batch_num, size_t nrank, size_t rank,
                   size_t cache_size, size_t num_threads);
/* TopkScheduler::start (laia/src/topk_scheduler.cc:47-187), the planner run_laia.py uses with
 * --local-shared: scores over the first top_k_table tables of the dataset's pre-profiled order
 * ("criteo" / "avazu" / "movie" / "criteosearch", :150-165), samples and slots split over
 * num_threads logical threads (the plan depends on that number, as in the reference).  Same
 * hb_laia_next / _plan / _dist surface afterwards. */
int hb_laia_create_topk(hb_laia **out, const uint64_t *sample_embs, size_t num_sample, size_t num_table,
                        size_t epoch_num, size_t mini_batch_size, size_t batch_num, size_t nrank,
                        size_t rank, size_t cache_size, size_t num_threads, const char *dataset,
                        size_t top_k_table);
int hb_laia_destroy(hb_laia *s);
/* Plan the next global batch (laia_scheduler.cc:115-169 one iteration: get_dist :171-271, then the
 * snapshot update :146-161).  *done = 1 when the sequence (epochs x batches, one more batch in the
 * last epoch) is over — the reference then pushes the terminator {0}. */
int hb_laia_next(hb_laia *s, int *done);
/* Results of the most recent hb_laia_next, for any worker (the reference hands out `rank`'s):
 * the communication plan (ascending unique keys) and the mini_batch_size sample positions. */
int hb_laia_plan_size(hb_laia *s, size_t worker, size_t *n);
int hb_laia_plan(hb_laia *s, size_t worker, uint64_t *keys, size_t cap);
int hb_laia_dist(hb_laia *s, size_t worker, uint64_t *sample_idx);
/* MiniLRUCache::get_keys of a worker's snapshot: valid keys, ascending (keys may be NULL: count) */
int hb_laia_snapshot_keys(hb_laia *s, size_t worker, uint64_t *keys, size_t cap, size_t *n);
/* Shared-memory message ring between the planning process of a node and its local workers — the
 * reference's SharedMemBuf (laia/include/share_mem.h:39-160, ring_buffer.h): POSIX shared memory
 * "laia_cache_<local rank>", single producer / single consumer, a message = length word + words.
 * send: *sent = n, or -1 when there is no room; recv: *n = length of the next message or -1 when
 * empty, consumed when data != NULL and cap >= length. */
typedef struct hb_shmring hb_shmring;
int hb_shmring_open(hb_shmring **out, const char *name, int create, size_t data_bytes);
int hb_shmring_close(hb_shmring *r);
int hb_shmring_send(hb_shmring *r, const uint64_t *data, size_t n, long long *sent);
int hb_shmring_recv(hb_shmring *r, uint64_t *data, size_t cap, long long *n);
int hb_shmring_used(hb_shmring *r, size_t *words);
/* The snapshot cache alone (laia/include/mini_lru_cache.h:14-137).  get returns -1 hit, -2 stale
 * hit, 0 miss, 1 miss that evicted a valid line (:69-105). */
typedef struct hb_minilru hb_minilru;
int hb_minilru_create(hb_minilru **out, size_t capacity);
int hb_minilru_destroy(hb_minilru *m);
int hb_minilru_get(hb_minilru *m, uint64_t key);
int hb_minilru_check(hb_minilru *m, uint64_t key);
int hb_minilru_outdate(hb_minilru *m, uint64_t key);
int hb_minilru_evict(hb_minilru *m, uint64_t key);
int hb_minilru_keys(hb_minilru *m, uint64_t *keys, size_t cap, size_t *n);

#ifdef __cplusplus
}
#endif
#endif /* HERALD_B200_H_ */
