set -x
mkdir -p gpurun_out
TAG=${TAG:-sp}
timeout 900 python -m pytest tests/test_split_reduce_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
for mode in exact split; do
HERALD_REDUCE=$mode timeout 600 python bench.py --config c5 --steps 20 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_c5_$mode.json 2> gpurun_out/${TAG}_c5_$mode.err
tail -3 gpurun_out/${TAG}_c5_$mode.err
python - gpurun_out/${TAG}_c5_$mode.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],4),'roofline',round(d['roofline']['frac'],3),{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})
PY
done
