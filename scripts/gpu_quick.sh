# quick GPU check: parity tests + bench variants.  usage: bash scripts/gpu_quick.sh TAG
set -x
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
HERALD_SEG_ROWS=2 timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_rows2.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json","gpurun_out/${TAG}_bench_rows2.json"):
    try:
        d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["phases"], d["roofline"]["kernels"])
    except Exception as e: print(f, "ERR", e)
PY
