set -x
mkdir -p gpurun_out
TAG=${TAG:-chk}
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e --seg-trace gpurun_out/${TAG}_segtrace.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_bench.json
