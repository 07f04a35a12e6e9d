set -x
mkdir -p gpurun_out
TAG=${1:-p}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for rows in 4 2; do
for thr in 64 32 16; do
  HERALD_SEG_ROWS=$rows HERALD_HOT_THRESHOLD=$thr timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_r${rows}_thr$thr.json 2>>gpurun_out/${TAG}.err
done; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_r*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["phases"]["push_ms"].items()}, {k:round(v,4) for k,v in d["phases"]["pull_ms"].items()})
    except Exception as e: print(f,"ERR",e)
PY
