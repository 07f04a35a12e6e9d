set -x
mkdir -p gpurun_out
TAG=${1:-mg}
G=${2:-2}
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --gpus $G --steps 30 --warmup 10 ${BENCH_ARGS} > gpurun_out/${TAG}_bench_g$G.json 2> gpurun_out/${TAG}_bench_g$G.err
tail -5 gpurun_out/${TAG}_bench_g$G.err
python - gpurun_out/${TAG}_bench_g$G.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step',round(d['ms_per_step'],4),'value',d['value'],'e2e',d.get('e2e'),'launches',d['gpu_launches'])
    print(d['phases'])
    print('nvlink',d['roofline'].get('nvlink'))
    print('parity',d.get('parity'))
except Exception as e: print('ERR',e)
PY
