# round 2, call A: parity of the warp-specialised segment reduce + first timings and variants
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2a}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
B="timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e"
$B --seg-trace gpurun_out/${TAG}_segtrace.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
HERALD_HOT_THRESHOLD=32 $B --seg-trace gpurun_out/${TAG}_segtrace_thr32.json > gpurun_out/${TAG}_bench_thr32.json 2>> gpurun_out/${TAG}_bench.err
HERALD_HOT_STAGES=4 $B --seg-trace gpurun_out/${TAG}_segtrace_st4.json > gpurun_out/${TAG}_bench_st4.json 2>> gpurun_out/${TAG}_bench.err
HERALD_TICKET_ROWS=16 $B > gpurun_out/${TAG}_bench_tk16.json 2>> gpurun_out/${TAG}_bench.err
HERALD_TICKET_ROWS=4 $B > gpurun_out/${TAG}_bench_tk4.json 2>> gpurun_out/${TAG}_bench.err
HERALD_MED_THRESHOLD=8 $B > gpurun_out/${TAG}_bench_med8.json 2>> gpurun_out/${TAG}_bench.err
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d['roofline']['kernels']['segment_reduce_kernel<AccumulatePush>']
    print(round(d['ms_per_step'],4), round(k['ms'],4), round(k['gbs']), d['phases']['push_ms'], d['phases']['pull_ms'])
except Exception as e: print('ERR',e)
PY
done
