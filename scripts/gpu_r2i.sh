# round 2: scale parity tests + default bench (with cpu baseline + parity record)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2i}
free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - gpurun_out/${TAG}_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],4),'e2e',d.get('e2e'),'launches',d['gpu_launches'])
    print('roofline',d['roofline']['kernel'],round(d['roofline']['frac'],3),{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})
    print(d['phases'])
    print('parity',d.get('parity'))
    print('cpu',d.get('cpu_baseline'))
except Exception as e: print('ERR',e)
PY
