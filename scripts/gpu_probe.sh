set -x
mkdir -p gpurun_out
TAG=${1:-p}
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2>>gpurun_out/${TAG}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["phases"]["push_ms"].items()}, {k:round(v,4) for k,v in d["phases"]["pull_ms"].items()})
    except Exception as e: print(f,"ERR",e)
PY
