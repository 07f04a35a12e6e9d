# round 2 iteration: parity tests + bench (device-resident + e2e) + launch list
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2h}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - gpurun_out/${TAG}_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],4),'e2e',d.get('e2e'),'launches',d['gpu_launches'])
    print('roofline',d['roofline']['kernel'],round(d['roofline']['frac'],3),{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})
    print(d['phases'])
except Exception as e: print('ERR',e)
PY
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/last_step.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_step.txt
fi
