# A/B of one environment knob on the default bench (device-resident numbers only), then the parity tests
# usage: KNOB=HERALD_ROW_TICKETS VALUES="0 1 0 1" [NO_TESTS=1] [BENCH_ARGS=...] bash scripts/gpu_ab.sh
set -x
mkdir -p gpurun_out
TAG=${TAG:-ab}
for V in ${VALUES}; do
  env ${KNOB}=$V timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-e2e --parity-steps ${PSTEPS:-0} ${BENCH_ARGS} > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err
  python - "gpurun_out/${TAG}_$V.json" "$V" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('AB',sys.argv[2],'ms/step',round(d['ms_per_step'],4),'launches',d['gpu_launches'],{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, {k:round(v*1e3,1) for k,v in d['phases']['pull_ms'].items()}, {k:round(v*1e3,1) for k,v in d['phases']['push_ms'].items()}, d.get('parity',{}).get('counters_equal'), d.get('parity',{}).get('rows_crc_equal'))
except Exception as e: print('ERR',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
if [ -z "$NO_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
fi
