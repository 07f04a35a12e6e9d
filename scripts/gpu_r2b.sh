# quick iteration: parity tests + bench with the segment_reduce timeline
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2b}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
B="timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e"
$B --seg-trace gpurun_out/${TAG}_segtrace.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
for v in $VARIANTS; do
  name=$(echo $v | tr '=,' '__')
  env $(echo $v | tr ',' ' ') $B --seg-trace gpurun_out/${TAG}_segtrace_$name.json > gpurun_out/${TAG}_bench_$name.json 2>> gpurun_out/${TAG}_bench.err
done
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d['roofline']['kernels']['segment_reduce_kernel<AccumulatePush>']
    print(round(d['ms_per_step'],4), round(k['ms'],4), round(k['gbs']), d['phases']['push_ms'], d['phases']['pull_ms'])
except Exception as e: print('ERR',e)
PY
done
for f in gpurun_out/${TAG}_segtrace*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('span',d['kernel_span_us'],'hot_end',[round(x,1) for x in d['cta_hot_end_us']],'end',[round(x,1) for x in d['cta_end_us']])
it=sorted(d['items'],key=lambda x:-x[2])[:6]
print('top items', [[round(x[0],1),round(x[1],1),x[2],round((x[1]-x[0])*1e3/x[2],2)] for x in it])
print('wait cyc/occ', [round(x,1) for x in d.get('item_wait_cycles_per_occurrence',[])[:48]])
print('ns/occ', [round(x,1) for x in d.get('item_ns_per_occurrence',[])[:48]])
PY
done
