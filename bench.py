#!/usr/bin/env python
"""bench.py — WDL-Criteo embedding step (BASELINE.json metric) on N B200s of one node.

One step = the hot path over one batch, in the order Hetu's BSP-prefetch loop issues it
(python/hetu/gpu_ops/ParameterServerCommunicate.py:48-52):
    embedding_update(batch t: ids [B,26], grads [B,26,D])      coalesce + SGD-folded push
    embedding_lookup(batch t+1: ids [B,26]) -> [B,26,D]         dedup, cache resolve, gather
through the reference-facing cache API (herald_b200.cstable.CacheSparseTable).

  value   samples/s with ids / grads / dest resident in HBM (device pointers), calls enqueued
          back to back, timed with CUDA events on the cache's stream, max over ranks
  e2e     the same step with HOST (pinned) buffers: ids + grads H2D and gathered rows D2H inside
          the timed region, `.wait()` after every call as the reference's loop does
  roofline   dominant kernel's algorithmic bytes / its CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference   the reference's own CPU cache + PS handler (oracle/_ref, the
          reference sources compiled unmodified; falls back to the C++ port) on this box's host

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIELDS = 26
VOCAB = 33762577          # examples/ctr/models/wdl_criteo.py:9
ZIPF_A = 1.05


# BASELINE.json `configs`, in order.  The per-GPU call shape is what a preset fixes; --gpus picks
# the group size (c3 / c4 are quoted on 8 GPUs, c5 is a sweep: scripts/sweep_c5.sh).
CONFIGS = {
    "c1": dict(batch=256, dim=128, vocab=VOCAB, policy="lru", bound=0, plan=False,
               note="wdl_criteo, 1 worker + local PS, LRU 0.1, batch 256, emb 128 (examples/ctr/run_hetu.py)"),
    "c2": dict(batch=8192, dim=128, vocab=VOCAB, policy="lru", bound=0, plan=False,
               note="wdl_criteo single B200, LRU 0.1, batch 8192, emb 128, 26 fields, Zipf(1.05)"),
    "c3": dict(batch=8192, dim=128, vocab=VOCAB, policy="lfu", bound=0, plan=False,
               note="dcn_criteo 8 GPUs row-sharded, LFU, bound 0 (BSP); same 26-field / V sparse side"),
    "c4": dict(batch=8192, dim=512, vocab=VOCAB, policy="lru", bound=10, plan=True,
               note="run_laia wdl_criteo 8 GPUs, Herald plans (update_with_push_keys), bound 10, emb 512"),
    "c5": dict(batch=65536, dim=128, vocab=100_000_000, policy="lru", bound=0, plan=False,
               note="scaling sweep point: 1e8-row table, emb 128/512 (--dim), batch 2k-64k (--batch)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="herald_b200", choices=["herald_b200", "reference"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=None)
    ap.add_argument("--policy", default=None)
    ap.add_argument("--bound", type=int, default=None)
    ap.add_argument("--ratio", type=float, default=0.1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--parity-steps", type=int, default=3,
                    help="steps at the start of the run whose counters / gathered rows / owner rows are "
                         "compared with the oracle (N = 1: on the benchmarked table; N > 1: whole-group "
                         "replay on a 1M-row side table); 0 = off")
    ap.add_argument("--ids", default="permuted", choices=["permuted", "folded"],
                    help="Zipf rank -> row: fixed permutation (default, SURVEY 8d) or rank == row")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="BASELINE.json configs[0..4] presets (c2 = the headline config, default); the "
                         "individual flags are filled from the preset unless given")
    ap.add_argument("--plan", action="store_true", default=None,
                    help="drive the updates with Laia / Herald communication plans "
                         "(embedding_update_with_push_keys, run_laia.py): c4")
    ap.add_argument("--seg-trace", default=None,
                    help="diagnostics: write the per-CTA timeline of one segment_reduce launch here")
    args = ap.parse_args()
    for k, v in CONFIGS[args.config].items():
        if k != "note" and getattr(args, k, None) is None:
            setattr(args, k, v)
    args.plan = bool(args.plan)
    global _IDS_MODE
    _IDS_MODE = args.ids
    return args


# ----------------------------------------------------------------------------------------------
# synthetic Criteo-shaped ids: unified Zipf(1.05) over the single [V, D] table, carried as float32
# exactly like Hetu's dataloader (python/hetu/dataloader.py:14) — ids above 2^24 round.
# ----------------------------------------------------------------------------------------------
_IDS_MODE = "permuted"
PERF_EVERY = 8          # phase events on every 8th step of the timed region


def _spread_multiplier(vocab):
    """A multiplier coprime with `vocab` near the golden-ratio point: rank -> (rank * m) % vocab is a
    fixed permutation of [0, vocab) that sends neighbouring popularity ranks far apart."""
    import math
    m = int(0.6180339887498949 * vocab) | 1
    while math.gcd(m, vocab) != 1:
        m += 2
    return m


def rank_to_id(ranks, vocab):
    """Popularity rank (0 = hottest) -> table row.  SURVEY 8(d) flavour (i): the Zipf rank is mapped
    through a fixed permutation of [0, V), as label-encoded Criteo ids are not sorted by
    popularity (examples/ctr/models/load_data.py:193-205 offsets 26 per-field encodings);
    `--ids folded` keeps rank == row (every hot row in the first shard)."""
    ranks = np.asarray(ranks, dtype=np.int64) % vocab
    if _IDS_MODE == "folded":
        return ranks
    return (ranks * _spread_multiplier(vocab)) % vocab


def make_ids(step, batch, vocab, rank=0):
    rng = np.random.default_rng(1234 + step + 100003 * rank)
    z = rng.zipf(ZIPF_A, (batch, FIELDS))
    return rank_to_id(z - 1, vocab).astype(np.float32)


def hottest_ids(lo, hi, vocab, dtype):
    """Rows of the popularity ranks [lo, hi) as the cache would see them (float32-carried ids round
    above 2^24: duplicates after rounding are dropped, order kept)."""
    ids = rank_to_id(np.arange(lo, hi), vocab).astype(np.float32)
    _, first = np.unique(ids, return_index=True)
    return ids[np.sort(first)].astype(dtype)


def rotating_buffers(args):
    """Gradient / destination buffer pairs the timed loop rotates through: at least 3, and enough
    that one round of them exceeds twice the 126 MB L2 (small batches: c1)."""
    pair = 2 * args.batch * FIELDS * args.dim * 4
    return int(min(64, max(3, -(-2 * 126_000_000 // pair))))


def laia_schedule(args, rank, world, nbatch):
    """Herald / Laia workload (run_laia.py): the planner — replicated on every rank, as in the
    reference — assigns the samples of each global batch to the workers and emits per worker the
    keys it must push so that its peers find them fresh in the NEXT batch (laia_scheduler.cc:115-271).
    -> (ids[b] float32 [B, 26] of this rank, plans[b] uint64 ascending) for b in [0, nbatch).
    plans[b + 1] is the push_keys argument of update(batch b) (laia_dataloader.py:108-114: the
    first plan is dropped).  Planning runs ahead of training in the reference (its own thread and a
    queue); here it is done before the timed region."""
    from herald_b200.laia import LaiaScheduler
    B, V = args.batch, args.vocab
    embs = np.concatenate([make_ids(b, B * world, V, 0).astype(np.uint64) for b in range(nbatch)])
    sched = LaiaScheduler()
    sched.start(embs, embs.shape[0], FIELDS, 1, B, nbatch, world, rank, cache_limit(V, args.ratio), 16)
    ids, plans = [], []
    t0 = time.perf_counter()
    for b in range(nbatch):
        if not sched.step():
            break
        plans.append(sched.plan_of(rank).copy())
        ids.append(embs[sched.dist_of(rank).astype(np.int64)].astype(np.float32))
    planner_ms = (time.perf_counter() - t0) * 1e3 / max(1, len(ids))
    sched.close()
    assert len(ids) == nbatch, "the planner ended early"
    return ids, plans, planner_ms


def cache_limit(vocab, ratio):
    return int(ratio * (vocab - 1)) + 1          # examples/ctr/run_hetu.py:256


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's CPU cache + in-process PS handler, driven synchronously
# ----------------------------------------------------------------------------------------------
def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def crc_rows(a):
    import zlib
    return zlib.crc32(np.ascontiguousarray(a, np.float32).view(np.uint8).reshape(-1))


COUNTERS = ("num_unique", "num_miss", "num_evict", "num_transfered")


def counters_of(perf_entry):
    return [int(perf_entry[k]) if k in perf_entry else 0 for k in COUNTERS]


def run_cpu_reference(args, steps, warmup, mirror=None):
    """-> dict(value, ms_per_step, kind, cores, sample, vocab[, parity]).  One pool thread does each
    call (ps-lite/src/thread_pool.cc:4 has 5 threads but a call runs on one; the cache PSFs are
    serial on the server: ps-lite/src/PSFhandle_embedding.cc:23-27).

    mirror (from the GPU arm, N = 1): {"touched", "rows0"} = the rows of the GPU's table this run
    will touch, as they were before the first call, and {"steps", "crc", "pull", "push",
    "keys_after", "rows_after", "ver_after"} = what the GPU arm observed on its first steps.  The
    oracle then starts from the same table, runs the same calls and the two are compared."""
    from oracle import port, ref
    kind = "reference" if ref.available() else "port"
    impl = ref if kind == "reference" else port
    vocab, D, B = args.vocab, args.dim, args.batch
    need_gb = vocab * D * 4 / 1e9 * 1.3 + 8
    folded = False
    if host_mem_available_gb() < need_gb:
        vocab, folded = 4_000_000, True
    limit = cache_limit(vocab, args.ratio)
    if mirror is not None and (folded or kind != "reference"):
        mirror = dict(skipped="host RAM too small for the full table" if folded else "oracle/_ref not built")
    if mirror is not None and "skipped" not in mirror:
        srv = impl.Server(vocab, D)                              # zeros; then the GPU table's rows
        srv.load_rows_at(mirror["touched"], mirror["rows0"])
        mirror["rows0"] = None
    elif kind == "reference":
        srv = impl.Server(vocab, D, init=(2, 0.0, 0.01, 123))   # Normal(0, 0.01): wdl_criteo.py:13-14
    else:
        srv = impl.Server(vocab, D)
    cache = impl.Cache(srv, args.policy, limit, args.bound)
    # fill the cache with the hottest ids (under (zipf-1) % V the small ids are the hot ones)
    t0 = time.perf_counter()
    chunk = 1 << 20
    for lo in range(0, limit, chunk):
        cache.embedding_lookup(hottest_ids(lo, min(lo + chunk, limit), vocab, np.uint64))
    fill_s = time.perf_counter() - t0
    N = B * FIELDS
    grads = make_grads(B, D, 0)
    dest = np.empty((N, D), np.float32)
    plans = None
    if args.plan:
        saved, args.vocab = args.vocab, vocab
        ids_f, plans, _ = laia_schedule(args, 0, 1, steps + warmup + 2)
        args.vocab = saved
        ids = [a.reshape(-1).astype(np.uint64) for a in ids_f]
    else:
        ids = [make_ids(s, B, vocab).reshape(-1).astype(np.uint64) for s in range(steps + warmup + 1)]
    cache.embedding_lookup(ids[0], dest)
    parity = None
    check = mirror is not None and "skipped" not in mirror
    P = mirror["steps"] if check else 0
    if check:
        parity = {"steps": P, "counters_equal": True, "rows_crc_equal": crc_rows(dest) == mirror["crc"][0]}
    elif mirror is not None:
        parity = {"steps": 0, "skipped": mirror["skipped"]}
    times = []
    for s in range(steps + warmup):
        t0 = time.perf_counter()
        cache.embedding_update(ids[s], grads, None if plans is None else plans[s + 1])
        cache.embedding_lookup(ids[s + 1], dest)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        if s < P:
            push, pull = counters_of(cache.perf[-2]), counters_of(cache.perf[-1])
            parity["counters_equal"] &= push == mirror["push"][s] and pull == mirror["pull"][s]
            parity["rows_crc_equal"] &= crc_rows(dest) == mirror["crc"][s + 1]
            if s == P - 1:
                rows, ver = srv.rows_at(mirror["keys_after"])
                parity["owner_rows_equal"] = bool(np.array_equal(rows.view(np.uint32),
                                                                 mirror["rows_after"].view(np.uint32)))
                parity["owner_versions_equal"] = bool(np.array_equal(ver, mirror["ver_after"]))
                parity["owner_rows_checked"] = int(rows.shape[0])
                parity["oracle"] = kind
                parity["scope"] = ("the benchmarked table and cache (V=%d, limit %d), pre-filled, first %d "
                                   "update+lookup steps: per-call counters, CRC32 of the gathered rows, owner "
                                   "rows + versions of every key touched, all bit-exact" % (vocab, limit, P))
    total = float(np.sum(times))
    sample = ("%d timed + %d warm-up steps of the same workload (B=%d, 26 fields, D=%d, V=%d%s, "
              "%s limit %d, bound %d) after filling the cache with the %d hottest ids (%.1f s); "
              "synchronous calls, in-process PS (no ZMQ/RDMA), median step %.1f ms" %
              (steps, warmup, B, D, vocab, " folded: host RAM" if folded else "", args.policy,
               limit, args.bound, limit, fill_s, 1e3 * float(np.median(times))))
    return dict(value=B * steps / total, ms_per_step=1e3 * total / steps, kind=kind, cores=1,
                sample=sample, vocab=vocab, parity=parity)


def make_grads(B, D, rank):
    """Gradient rows of one batch, already multiplied by -lr (ParameterServerCommunicate.py:58-59)."""
    rng = np.random.default_rng(7 + rank)
    return (rng.normal(0, 1e-3, (B * FIELDS, D)) * 1e-2).astype(np.float32)


def reference_main(args, rank, world):
    if rank != 0:
        return
    res = run_cpu_reference(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "wdl_criteo_embedding_step_samples_per_sec",
        "value": res["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, res["vocab"]),
        "cpu_baseline": {"value": res["value"], "unit": "samples/s", "cores": res["cores"],
                         "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, vocab=None):
    return {"workload": "%s: embedding step = update(batch t)%s + lookup(batch t+1), "
                        "%s cache ratio %.2f, bound %d" % (
                            args.config, " with Herald plans (push_keys)" if args.plan else "",
                            args.policy.upper(), args.ratio, args.bound),
            "preset": CONFIGS[args.config]["note"],
            "batch_per_gpu": args.batch, "fields": FIELDS, "emb_dim": args.dim,
            "table_rows": vocab or args.vocab, "cache_limit": cache_limit(vocab or args.vocab, args.ratio),
            "ids": "Zipf(%.2f) unified, %s, float32-carried" % (ZIPF_A, "rank -> row by a fixed permutation" if _IDS_MODE == "permuted" else "rank == row"),
            "l2": "inputs larger than L2: %d rotating grads/dest buffer pairs of 2x%.1f MB, "
                  "%.0f GB table" % (rotating_buffers(args), args.batch * FIELDS * args.dim * 4 / 1e6,
                                     (vocab or args.vocab) * args.dim * 4 / 1e9),
            "parallelism": "dp%d row-sharded" % args.gpus}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def dl_stream_of(cuda_stream_value):
    from herald_b200.stream import DLStream
    holder = ctypes.c_void_p(cuda_stream_value)
    st = DLStream()
    st.device_id = 0
    st.handle = ctypes.cast(ctypes.pointer(holder), ctypes.c_void_p)

    class _S(object):
        pass

    s = _S()
    s.handle = ctypes.pointer(st)
    s._keep = (holder, st)
    return s


def seg_trace(path, run_step):
    """Diagnostics (untimed): per-CTA / per-hot-item timeline of one segment_reduce launch."""
    from herald_b200._base import _LIB, check_call
    words = 2 + 4 * 2048 + 3 * 1024
    check_call(_LIB.HBSegTraceEnable(1))
    run_step()
    buf = (ctypes.c_ulonglong * words)()
    check_call(_LIB.HBSegTraceRead(buf, ctypes.c_size_t(words)))
    check_call(_LIB.HBSegTraceEnable(0))
    a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
    grid, nitems = int(a[0]), int(a[1])
    cta = a[2:2 + 4 * 2048].reshape(2048, 4)[:grid]
    items = a[2 + 4 * 2048:].reshape(1024, 3)[:min(nitems, 1024)].copy()
    waited = (items[:, 2] >> 32) * 16          # adder cycles spent waiting for the ring
    items[:, 2] &= 0xffffffff
    t0 = int(cta[:, 0].min())
    out = {"grid": grid, "hot_items": nitems,
           "kernel_span_us": (int(cta[:, 2].max()) - t0) / 1e3,
           "cta_start_us": np.percentile((cta[:, 0] - t0) / 1e3, [0, 50, 100]).tolist(),
           "cta_hot_end_us": np.percentile((cta[:, 1] - t0) / 1e3, [0, 10, 50, 90, 100]).tolist(),
           "cta_end_us": np.percentile((cta[:, 2] - t0) / 1e3, [0, 10, 50, 90, 100]).tolist(),
           "cta_items": np.percentile(cta[:, 3], [0, 50, 100]).tolist(),
           "items": [[(int(x[0]) - t0) / 1e3, (int(x[1]) - t0) / 1e3, int(x[2])] for x in items[:64]],
           "item_ns_per_occurrence": [float((x[1] - x[0]) / max(1, x[2])) for x in items[:64]],
           "item_wait_cycles_per_occurrence": [float(w / max(1, x[2])) for w, x in zip(waited[:64], items[:64])],
           "last_item_end_us": (int(items[:, 1].max()) - t0) / 1e3 if len(items) else 0.0}
    with open(path, "w") as f:
        json.dump(out, f)


def group_parity(args, rank, world, local_rank, dist, comm):
    """N > 1: whole-group parity run BEFORE the benchmark, on a side table small enough to mirror
    on the host (1 000 003 rows) but with the benchmark's call shape (B x 26 keys per rank, D,
    policy, bound) — so full-size batches travel through the owner mailboxes and the remote pulls.
    Every rank drives its cache for P update+lookup steps; rank 0 replays the WHOLE group on the
    oracle (one server, one cache per rank, calls in rank order: the order the owners apply the
    mailboxes in) and compares every rank's counters, CRC32 of gathered rows, and final shard
    rows + versions.  Returns the parity record on rank 0, None elsewhere."""
    import herald_b200 as hb
    from herald_b200 import ps, partition
    from herald_b200.cstable import CacheSparseTable
    B, D, P = args.batch, args.dim, args.parity_steps
    V2, limit2, node = 1_000_003, 300_000, 7
    r = np.arange(V2, dtype=np.int64)[:, None]
    c = np.arange(D, dtype=np.int64)[None, :]
    rows = (((r * 31 + c * 17) % 2003) - 1001).astype(np.float32) * np.float32(1e-5)
    table = comm.InitTensor(node, ps.kCacheTable, V2, D, ps.Constant, 0.0)
    table.load_rows(rows)                                      # clipped to this rank's shard
    cst = CacheSparseTable(limit2, V2, D, node, args.policy, args.bound)
    cst.cache.reserve(max(B * FIELDS, 1 << 20))
    cst.perf_enabled(True)
    comm.BarrierWorker()
    host = hb.cpu(0)
    ids = [[make_ids(5000 + s, B, V2, w) for s in range(P + 1)] for w in range(world)]
    mine = {"crc": [], "pull": [], "push": []}
    g_host, d_host = hb.array(make_grads(B, D, rank).reshape(B, FIELDS, D), host), hb.empty((B, FIELDS, D), host)
    cst.embedding_lookup(hb.array(ids[rank][0], host), d_host, sync=True)
    mine["crc"].append(crc_rows(d_host.host_view()))
    for s in range(P):
        cst.embedding_update(hb.array(ids[rank][s], host), g_host, sync=True)
        comm.BarrierWorker()                                   # BSP: every push before any pull
        cst.embedding_lookup(hb.array(ids[rank][s + 1], host), d_host, sync=True)
        mine["push"].append(counters_of(cst.perf[-2]))
        mine["pull"].append(counters_of(cst.perf[-1]))
        mine["crc"].append(crc_rows(d_host.host_view()))
    comm.BarrierWorker()
    import zlib
    mine["shard_rows_crc"] = crc_rows(table.read_rows())
    mine["shard_ver_crc"] = zlib.crc32(np.ascontiguousarray(table.read_versions()).view(np.uint8))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    comm.BarrierWorker()
    del cst, g_host, d_host
    comm.ClearTensor(node)
    if rank != 0:
        return None
    from oracle import port, ref
    kind = "reference" if ref.available() else "port"
    impl = ref if kind == "reference" else port
    srv = impl.Server(V2, D, rows)
    caches = [impl.Cache(srv, args.policy, limit2, args.bound) for _ in range(world)]
    grads = [make_grads(B, D, w) for w in range(world)]
    rec = {"steps": P, "ranks": world, "counters_equal": True, "rows_crc_equal": True}
    dest = np.empty((B * FIELDS, D), np.float32)

    def lookups(s):
        for w in range(world):
            caches[w].embedding_lookup(ids[w][s].reshape(-1).astype(np.uint64), dest)
            rec["rows_crc_equal"] &= crc_rows(dest) == everyone[w]["crc"][s]
            if s:
                rec["counters_equal"] &= counters_of(caches[w].perf[-1]) == everyone[w]["pull"][s - 1]

    lookups(0)
    for s in range(P):
        for w in range(world):
            caches[w].embedding_update(ids[w][s].reshape(-1).astype(np.uint64), grads[w])
            rec["counters_equal"] &= counters_of(caches[w].perf[-1]) == everyone[w]["push"][s]
        lookups(s + 1)
    orows, over = srv.rows(), srv.versions()
    ok_rows = ok_ver = True
    for w in range(world):
        b, n = partition.shard_range(w, world, V2)
        ok_rows &= crc_rows(orows[b:b + n]) == everyone[w]["shard_rows_crc"]
        ok_ver &= zlib.crc32(np.ascontiguousarray(over[b:b + n]).view(np.uint8)) == everyone[w]["shard_ver_crc"]
    rec["owner_rows_equal"], rec["owner_versions_equal"] = bool(ok_rows), bool(ok_ver)
    rec["oracle"] = kind
    rec["scope"] = ("whole group of %d ranks on a %d-row side table (limit %d), benchmark call shape "
                    "(%d keys per rank and call, D=%d, %s, bound %d), %d update+lookup steps replayed on the "
                    "oracle in rank order: every rank's counters, CRC32 of gathered rows, final shard rows "
                    "and versions, bit-exact" % (world, V2, limit2, B * FIELDS, D, args.policy, args.bound, P))
    return rec


def herald_main(args, rank, world, local_rank):
    import herald_b200 as hb
    from herald_b200 import ps, stream as hstream
    from herald_b200._base import _LIB, check_call, kernel_launch_count
    from herald_b200.cstable import CacheSparseTable

    dist = None
    if world > 1:
        import torch.distributed as dist                       # rendezvous + barrier only (gloo)
        dist.init_process_group("gloo", rank=rank, world_size=world)

        def exchange(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]

        # NCCL prints "NCCL version ..." on file descriptor 1 while the communicator comes up:
        # keep stdout for the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            ps.group_init(rank, world, local_rank, exchange)
        finally:
            sys.stdout.flush()
            try:
                ctypes.CDLL(None).fflush(None)      # anything still in C stdio buffers goes to stderr too
            except Exception:
                pass
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    dev = hb.gpu(local_rank)
    B, D, V = args.batch, args.dim, args.vocab
    N = B * FIELDS
    limit = cache_limit(V, args.ratio)
    comm = hb.worker_init(local_rank)
    group_par = None
    if world > 1 and args.parity_steps > 0:
        group_par = group_parity(args, rank, world, local_rank, dist, comm)
    table = comm.InitTensor(0, ps.kCacheTable, V, D, ps.Normal, 0.0, 0.01, 123)
    cst = CacheSparseTable(limit, V, D, 0, args.policy, args.bound)
    cst.cache.reserve(max(N, 1 << 20))
    stream = dl_stream_of(cst.cache.stream)
    ev = [hstream.create_event_handle(dev) for _ in range(4)]

    # ---- inputs ----
    K, W = args.steps, args.warmup
    # N = 1: the first P steps are the parity steps (compared with the oracle in the cpu_baseline
    # leg, which replays them from the same table rows); warm-up and timed steps follow
    P = args.parity_steps if (world == 1 and not args.no_cpu_baseline) else 0
    cpu_total = args.cpu_steps + 2
    P = min(P, cpu_total)
    total = P + K + W + 1
    plans_np, planner_ms = None, None
    if args.plan:
        P = 0                                      # (parity of the plan path: tests/test_cache_gpu.py, test_scale_gpu.py)
        ids_np, plans_np, planner_ms = laia_schedule(args, rank, world, total + 1)
    else:
        ids_np = [make_ids(s, B, V, rank) for s in range(max(total, cpu_total + 1))]
    ids_dev = [hb.array(a, dev) for a in ids_np[:total]]
    R = rotating_buffers(args)
    grads_np = make_grads(B, D, rank).reshape(B, FIELDS, D)                      # already x(-lr)
    grads_dev = [hb.array(grads_np, dev) for _ in range(R)]
    dest_dev = [hb.empty((B, FIELDS, D), dev) for _ in range(R)]
    mirror = None
    if P:
        # rows the oracle's run will touch, as they are before the first call
        chunk = 1 << 20
        fill = [hottest_ids(lo, min(lo + chunk, limit), V, np.uint64) for lo in range(0, limit, chunk)]
        touched = np.unique(np.concatenate(fill + [a.reshape(-1).astype(np.uint64)
                                                   for a in ids_np[:cpu_total + 1]]))
        mirror = {"steps": P, "touched": touched, "rows0": table.read_rows_at(touched)[0],
                  "crc": [], "pull": [], "push": []}
        del fill

    # ---- fill the cache with the hottest ids (setup, untimed) ----
    chunk = 1 << 20
    for lo in range(0, limit, chunk):
        n = min(chunk, limit - lo)
        k = hb.array(hottest_ids(lo, lo + n, V, np.float32), dev)
        d = hb.empty((k.shape[0], D), dev)
        cst.embedding_lookup(k, d, sync=True)
        del k, d
    if P:
        # parity steps: synchronous calls with host (pinned) buffers, everything observable recorded
        host = hb.cpu(0)
        g_host, d_host = hb.array(grads_np, host), hb.empty((B, FIELDS, D), host)
        cst.perf_enabled(True)
        cst.embedding_lookup(hb.array(ids_np[0], host), d_host, sync=True)
        mirror["crc"].append(crc_rows(d_host.host_view()))
        for s in range(P):
            cst.embedding_update(hb.array(ids_np[s], host), g_host, sync=True)
            cst.embedding_lookup(hb.array(ids_np[s + 1], host), d_host, sync=True)
            mirror["push"].append(counters_of(cst.perf[-2]))
            mirror["pull"].append(counters_of(cst.perf[-1]))
            mirror["crc"].append(crc_rows(d_host.host_view()))
        mirror["keys_after"] = np.unique(np.concatenate([a.reshape(-1).astype(np.uint64)
                                                         for a in ids_np[:P + 1]]))
        mirror["rows_after"], mirror["ver_after"] = table.read_rows_at(mirror["keys_after"])
        cst.perf_enabled(False)
        del g_host, d_host
    else:
        cst.embedding_lookup(ids_dev[0], dest_dev[0], sync=True)

    def step(s, keys, grads, dests, sync):
        if plans_np is not None:
            w1 = cst.embedding_update_with_push_keys(keys[s], plans_np[s + 1], grads[s % len(grads)], sync=sync)
        else:
            w1 = cst.embedding_update(keys[s], grads[s % len(grads)], sync=sync)
        w2 = cst.embedding_lookup(keys[s + 1], dests[s % len(dests)], sync=sync)
        return w1, w2

    def barrier():
        check_call(_LIB.DLStreamSync(stream.handle))
        if dist is not None:
            dist.barrier()

    # ---- warm-up (untimed) ----
    for s in range(P, P + W):
        step(s, ids_dev, grads_dev, dest_dev, False)
    if not os.environ.get("HB_BENCH_NOPERF"):       # diagnostics: cost of the phase events
        cst.perf_enabled(True)
        # CUDA events around the kernels of every 8th step (an event between two kernels costs
        # their launch overlap: 8 % of the step when every call carries them)
        cst.cache.set_perf_sampling(PERF_EVERY)
    barrier()

    # ---- timed region: K steps, device-resident inputs ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = kernel_launch_count()
    barrier()
    ev[0].record(stream)
    last = None
    for s in range(P + W, P + W + K):
        last = step(s, ids_dev, grads_dev, dest_dev, False)
    ev[1].record(stream)
    last[1].wait()
    barrier()
    launches = kernel_launch_count() - launches0
    ms = ev[1].time_since(ev[0])
    if os.environ.get("HB_BENCH_NOPERF"):
        if rank == 0:
            print(json.dumps({"diag": "phase events off", "ms_per_step": ms / K}), flush=True)
        os._exit(0)
    perf = list(cst.perf)[-2 * K:]

    if args.seg_trace and rank == 0:
        seg_trace(args.seg_trace, lambda: step(P + W + K - 1, ids_dev, grads_dev, dest_dev, True))

    # ---- end to end with host buffers (pinned NDArrays: the reference's calling convention) ----
    e2e = None
    if not args.no_e2e:
        cst.perf_enabled(False)
        Ke = min(K, 20)
        host = hb.cpu(0)
        g0 = P + W                                  # global index of the first e2e step (plans stay aligned)
        ids_host = {g0 + s: hb.array(ids_np[g0 + s], host) for s in range(Ke + 1)}
        grads_host = [hb.array(grads_np, host) for _ in range(2)]
        dest_host = [hb.empty((B, FIELDS, D), host) for _ in range(2)]
        cst.embedding_lookup(ids_host[g0], dest_host[0], sync=True)
        step(g0, ids_host, grads_host, dest_host, True)         # warm the staging buffers
        barrier()
        # Software-pipelined like the reference's prefetch loop (ParameterServerCommunicate.py:48-52:
        # the push is waited for, the pull is only waited for when its rows are consumed): the
        # update is waited for at once, lookup(t+1)'s rows are read on the host one step later, so
        # their download overlaps the upload of the next step's gradients (PCIe is full duplex).
        checksum = 0.0
        prev = None
        t_host0 = time.perf_counter()
        ev[2].record(stream)
        for s in range(g0 + 1, g0 + Ke):
            w1, w2 = step(s, ids_host, grads_host, dest_host, False)
            w1.wait()
            if prev is not None:
                prev[0].wait()
                checksum += float(prev[1].host_view()[0, 0, 0])   # the result is read on the host
            prev = (w2, dest_host[s % len(dest_host)])
        prev[0].wait()
        checksum += float(prev[1].host_view()[0, 0, 0])
        ev[3].record(stream)
        barrier()
        t_host = (time.perf_counter() - t_host0) * 1e3
        e2e_ms = max(ev[3].time_since(ev[2]), t_host)
        e2e = {"ms": e2e_ms, "steps": Ke - 1, "checksum": checksum}

    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (device-resident, e2e)

    # ---- max over ranks ----
    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e["ms"] / e2e["steps"] if e2e else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if e2e:
            e2e["ms"] = float(t[1]) * e2e["steps"]

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        pulls = [p for p in perf if p["type"] == "Pull"]
        pushes = [p for p in perf if p["type"] == "Push"]
        timed_pulls = [p for p in pulls if p["copy_time"] > 0] or pulls
        timed_pushes = [p for p in pushes if p["copy_time"] > 0] or pushes
        U_pull = float(np.mean([p["num_unique"] for p in pulls]))
        U_push = float(np.mean([p["num_unique"] for p in pushes]))
        t_gather = float(np.mean([p["copy_time"] for p in timed_pulls]))   # ms, gather kernel
        # the accumulate+push kernel alone (events right around segment_reduce_kernel; copy_time also
        # holds its plan kernel, reported as seg_plan_ms)
        t_accum = float(np.mean([p.get("kernel_time") or p["copy_time"] for p in timed_pushes]))
        t_plan = float(np.mean([p["copy_time"] for p in timed_pushes])) - t_accum
        row = D * 4
        gather_bytes = (U_pull + N) * row
        accum_bytes = (N + 2 * U_push) * row                              # SURVEY §8(d), bound 0
        pushed = float(np.mean([p["num_transfered"] - p["num_evict"] for p in pushes]))
        accum_bytes_owner = accum_bytes + 2 * pushed * row                # + owner row RMW (a9)
        kernels = {
            "gather_rows_kernel": {"ms": t_gather, "algorithmic_bytes": gather_bytes,
                                   "gbs": gather_bytes / t_gather / 1e6 if t_gather else None},
            "segment_reduce_kernel<AccumulatePush>": {
                "ms": t_accum, "seg_plan_ms": t_plan, "algorithmic_bytes": accum_bytes,
                "gbs": accum_bytes / t_accum / 1e6 if t_accum else None,
                "algorithmic_bytes_with_owner_row_rmw": accum_bytes_owner,
                "gbs_with_owner_row_rmw": accum_bytes_owner / t_accum / 1e6 if t_accum else None},
        }
        dom = max(kernels, key=lambda k: kernels[k]["ms"] or 0.0)
        achieved = kernels[dom]["gbs"] or 0.0
        phase = {
            "sampled_steps": len(timed_pushes),
            "pull_ms": {k: float(np.mean([p[k] for p in timed_pulls])) for k in
                        ("time", "sort_time", "lookup_time", "transfer_time", "copy_time", "insert_time")},
            "push_ms": {k: float(np.mean([p[k] for p in timed_pushes])) for k in
                        ("time", "sort_time", "lookup_time", "copy_time", "transfer_time")},
            "unique_per_lookup": U_pull, "miss_per_lookup": float(np.mean([p["num_miss"] for p in pulls])),
            "rows_pulled_per_lookup": float(np.mean([p["num_transfered"] for p in pulls])),
            "lines_pushed_per_update": float(np.mean([p["num_transfered"] for p in pushes])),
        }
        line = {
            "metric": "wdl_criteo_embedding_step_samples_per_sec",
            "value": world * B * K / (ms / 1e3), "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(dom),
                         "peak_source": peak_src, "kernels": kernels},
            "phases": phase,
        }
        if planner_ms is not None:
            line["planner"] = {"ms_per_global_batch": planner_ms, "threads": 16,
                               "plan_keys_per_update": float(np.mean([len(p) for p in plans_np])),
                               "note": "host planner (csrc/hb_laia.cu), run before the timed region"}
        if world > 1:
            # NVLink 5: 900 GB/s per direction per GPU (nominal).  Pull = rows this rank's sync kernel
            # reads out of peers' shards (inbound); push = lines its accumulate kernel deposits in
            # peers' mailboxes (outbound).  Bytes are the algorithmic row bytes, times are the CUDA-event
            # times of the kernels that move them (which also do their local work in that time).
            rp = float(np.mean([p.get("num_remote", 0) for p in pulls])) * row
            rq = float(np.mean([p.get("num_remote", 0) for p in pushes])) * row
            t_sync = float(np.mean([p["transfer_time"] for p in timed_pulls]))
            pull_gbs = rp / t_sync / 1e6 if t_sync else None
            push_gbs = rq / t_accum / 1e6 if t_accum else None
            line["roofline"]["nvlink"] = {
                "peak_gbs_per_direction": 900.0,
                "pull": {"kernel": "sync_kernel", "remote_bytes": rp, "ms": t_sync, "gbs": pull_gbs,
                         "frac_of_900": pull_gbs / 900.0 if pull_gbs else None},
                "push": {"kernel": "segment_reduce_kernel<AccumulatePush>", "remote_bytes": rq,
                         "ms": t_accum, "gbs": push_gbs,
                         "frac_of_900": push_gbs / 900.0 if push_gbs else None},
                "step": {"remote_bytes_in_plus_out": rp + rq, "gbs_over_whole_step": (rp + rq) / (ms / K) / 1e6}}
        if e2e:
            line["e2e"] = {"value": world * B * e2e["steps"] / (e2e["ms"] / 1e3), "unit": "samples/s",
                           "h2d_bytes_per_step": 2 * N * 4 + N * D * 4, "d2h_bytes_per_step": N * D * 4,
                           "ms_per_step": e2e["ms"] / e2e["steps"]}
        if world == 1 and not args.no_cpu_baseline:
            cpu = run_cpu_reference(args, args.cpu_steps, 2, mirror)
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "samples/s", "cores": cpu["cores"],
                                    "kind": cpu["kind"], "sample": cpu["sample"]}
            if cpu["parity"] is not None:
                line["parity"] = cpu["parity"]
        if group_par is not None:
            line["parity"] = group_par
        print(json.dumps(line), flush=True)

    del cst
    comm.ClearTensor(0)
    if dist is not None:
        ps.group_finalize()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_main(args, rank, world)
    else:
        herald_main(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
